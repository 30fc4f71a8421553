#!/bin/bash
# thread pass with lane-level refill: regression (bit-exact step bounds) + A/B bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02w_pytest.log 2>&1; tail -4 gpurun_out/r02w_pytest.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_$name.json 2> gpurun_out/r02w_bench_$name.err
  tail -c 200 gpurun_out/r02w_bench_$name.err | grep -v PARITY
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02w_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "narrow", round(s["ccd_narrow"],4), "parity", d["parity"] and d["parity"]["ok"], d["config"]["full_ccd_candidates_survivors_warnings_deferred_boxesThreadPass_boxesWarpPass_longestPairCycles_totalCycles"][2:])
PY
}
run refill IPCGPU_TI_REFILL=1
run norefill IPCGPU_TI_REFILL=0
run refill_b32 IPCGPU_TI_REFILL=1 IPCGPU_TI_BUDGET=32
run refill_b48 IPCGPU_TI_REFILL=1 IPCGPU_TI_BUDGET=48
