#!/bin/bash
# cell-centric EE pair kernel + side-stream pair-Hessian projection + fused fetch collective: regression tests, then A/B bench lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02j_pytest.log 2>&1
tail -6 gpurun_out/r02j_pytest.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_$name.json 2> gpurun_out/r02j_bench_$name.err
  tail -c 200 gpurun_out/r02j_bench_$name.err | grep -v PARITY
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02j_bench_$name.json").read().strip().splitlines()[-1])
print("$name value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
}
run new IPCGPU_PAIRS_MODE=1 IPCGPU_BARRIER_OVERLAP=1
run oldpairs IPCGPU_PAIRS_MODE=0 IPCGPU_BARRIER_OVERLAP=1
run nooverlap IPCGPU_PAIRS_MODE=1 IPCGPU_BARRIER_OVERLAP=0
