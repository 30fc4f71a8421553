#!/bin/bash
# tet kernel compiled for 6 / 8 / 10 / 12 CTAs per SM (register cap 156 / 128 / 96 / 80): stage time of the kernel and whole iteration
mkdir -p gpurun_out
for m in 8 6 10 12; do
IPCGPU_TET_MINB=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench_$m.json 2> gpurun_out/r02aa_bench_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02aa_bench_$m.json").read().strip().splitlines()[-1])
print("minb $m value", round(d["value"],4), "tet", round(d["stage_ms"]["elastic_tet"],4), "parity", d["parity"] and d["parity"]["ok"])
PY
done
