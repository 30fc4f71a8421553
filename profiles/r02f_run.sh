#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02f_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],3), "narrow", round(d["stage_ms"]["ccd_narrow"],3), "asm", round(d["stage_ms"]["assemble_csr"],3), d["parity"]["ok"] if d["parity"] else None, d["config"]["full_ccd_candidates_survivors_warnings_deferred_boxesThreadPass_boxesWarpPass_longestPairCycles_totalCycles"])
PY
}
run m0 IPCGPU_TI_MODE=0
run m0_occ3 IPCGPU_TI_MODE=0 IPCGPU_TI_OCC=3
run m0_b24 IPCGPU_TI_MODE=0 IPCGPU_TI_BUDGET=24
run m3 IPCGPU_TI_MODE=3
run m3_occ3 IPCGPU_TI_MODE=3 IPCGPU_TI_OCC=3
run m3_b64 IPCGPU_TI_MODE=3 IPCGPU_TI_BUDGET2=64
