#!/bin/bash
# level-buffer size of the Tight-Inclusion thread pass (local memory footprint vs deferrals); row-run lookups of the cell kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_contact.py tests/test_gpu_ccd.py tests/test_gpu_deferred.py -x -q ) 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02s_bench_$name.json 2> gpurun_out/r02s_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02s_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "narrow", round(s["ccd_narrow"],4), "broad", round(s["ccd_broad"],4), "cs", round(s["constraint_set"],4), d["config"]["full_ccd_candidates_survivors_warnings_deferred_boxesThreadPass_boxesWarpPass_longestPairCycles_totalCycles"][3:])
PY
}
run cap12
run cap8 IPCGPU_TI_CAP=8
run cap6 IPCGPU_TI_CAP=6
run cap16 IPCGPU_TI_CAP=16
run cap8_b32 IPCGPU_TI_CAP=8 IPCGPU_TI_BUDGET=32
run cap8_b16 IPCGPU_TI_CAP=8 IPCGPU_TI_BUDGET=16
run smem12 IPCGPU_TI_LVL_SMEM=1
run smem8 IPCGPU_TI_LVL_SMEM=1 IPCGPU_TI_CAP=8
run smem12_b32 IPCGPU_TI_LVL_SMEM=1 IPCGPU_TI_BUDGET=32
