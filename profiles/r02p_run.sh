#!/bin/bash
# fused energy + 32-bit grid keys: regression; then Tight-Inclusion thread-pass budget sweep after the code-size refactor
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02p_pytest.log 2>&1; tail -4 gpurun_out/r02p_pytest.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02p_bench_$name.json 2> gpurun_out/r02p_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02p_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "narrow", round(s["ccd_narrow"],4), "hash", round(s["hash"],4), "E", round(s["elastic_energy"],4), "tet", round(s["elastic_tet"],4), d["config"]["full_ccd_candidates_survivors_warnings_deferred_boxesThreadPass_boxesWarpPass_longestPairCycles_totalCycles"][3:])
PY
}
run b10
run b16 IPCGPU_TI_BUDGET=16
run b24 IPCGPU_TI_BUDGET=24
run b32 IPCGPU_TI_BUDGET=32
run b64 IPCGPU_TI_BUDGET=64
run b128 IPCGPU_TI_BUDGET=128
run m3_b32_b2_1024 IPCGPU_TI_MODE=3 IPCGPU_TI_BUDGET=32 IPCGPU_TI_BUDGET2=1024
