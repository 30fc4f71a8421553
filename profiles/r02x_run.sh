#!/bin/bash
# batched lane refill of the thread pass
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_deferred.py -x -q ) 2>&1 | tail -2
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02x_bench_$name.json 2> gpurun_out/r02x_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02x_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "narrow", round(s["ccd_narrow"],4))
PY
}
run batch1 IPCGPU_TI_REFILL_BATCH=1
run batch8 IPCGPU_TI_REFILL_BATCH=8
run batch12 IPCGPU_TI_REFILL_BATCH=12
run batch16 IPCGPU_TI_REFILL_BATCH=16
run batch24 IPCGPU_TI_REFILL_BATCH=24
run batch32 IPCGPU_TI_REFILL_BATCH=32
