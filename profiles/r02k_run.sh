#!/bin/bash
# CUDA-graph replay of the iteration: test, then bench lines graph / eager / graph + side-stream overlap
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_ccd.py -x -q ) > gpurun_out/r02k_pytest.log 2>&1
tail -12 gpurun_out/r02k_pytest.log
run() { name=$1; shift
  env "$@" > gpurun_out/r02k_bench_$name.json 2> gpurun_out/r02k_bench_$name.err
  tail -c 300 gpurun_out/r02k_bench_$name.err | grep -v PARITY
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_$name.json").read().strip().splitlines()[-1])
print("$name value", round(d["value"],4), "eager_profiled", round(d["config"]["eager_profiled_ms_per_step"],4), "e2e", round(d["e2e"]["value"],4), "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
}
run graph timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run eager timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --eager
run graph_overlap IPCGPU_BARRIER_OVERLAP=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
