#!/bin/bash
# graph replay with / without the side-stream overlap; ncu source-level capture of the warp-level Tight-Inclusion pass (full CCD instance)
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" > gpurun_out/r02l_bench_$name.json 2> gpurun_out/r02l_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02l_bench_$name.json").read().strip().splitlines()[-1])
print("$name value", round(d["value"],4), "eager_profiled", round(d["config"]["eager_profiled_ms_per_step"],4), "e2e", round(d["e2e"]["value"],4), "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
}
run graph timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run graph_overlap IPCGPU_BARRIER_OVERLAP=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --eager"
for k in k_ti_stage2 k_ti_stage15; do
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:^${k}" -s 9 -c 1 -f -o gpurun_out/r02l_prof_${k} $B > gpurun_out/r02l_prof_${k}.log 2>&1
tail -3 gpurun_out/r02l_prof_${k}.log
python profiles/summarize.py gpurun_out/r02l_prof_${k}.ncu-rep > gpurun_out/r02l_prof_${k}.summary.csv 2>/dev/null
ncu -i gpurun_out/r02l_prof_${k}.ncu-rep --page source --csv > gpurun_out/r02l_prof_${k}.source.csv 2>/dev/null
done
ls -la gpurun_out | grep r02l
