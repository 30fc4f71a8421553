#!/bin/bash
# parameter sweeps: tiles per CTA of the per-tet kernel, occupancy / budget of the Tight-Inclusion thread pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elastic.py -x -q 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02o_bench_$name.json 2> gpurun_out/r02o_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02o_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "tet", round(s["elastic_tet"],4), "asm", round(s["assemble_csr"],4), "narrow", round(s["ccd_narrow"],4), "frac", round(d["roofline"]["frac"],3))
PY
}
run tiles1 IPCGPU_TET_TILES=1
run tiles2 IPCGPU_TET_TILES=2
run tiles4 IPCGPU_TET_TILES=4
run occ3 IPCGPU_TI_OCC=3
run occ3_b6 IPCGPU_TI_OCC=3 IPCGPU_TI_BUDGET=6
run b6 IPCGPU_TI_BUDGET=6
run b16 IPCGPU_TI_BUDGET=16
