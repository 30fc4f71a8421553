#!/bin/bash
# state of the tree at the end of round 2: full GPU test suite, the bench line as the driver runs it, launch list, ncu captures of the pair kernels
# (the two kernels whose hit staging changed after the last capture)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02af_pytest.log 2>&1; tail -4 gpurun_out/r02af_pytest.log
timeout 900 python bench.py > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err || tail -5 gpurun_out/r02af_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02af_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "parity", d["parity"] and d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], "roofline", d["roofline"]["frac"], "launches", d["gpu_launches"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launch_bench.log 2>&1
python profiles/launches_by_kernel.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_by_kernel.csv
head -24 gpurun_out/r02_launches_by_kernel.csv
for k in k_cell_pairs_ee k_cell_pairs_pt; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 5 -c 1 -f -o gpurun_out/r02_prof_${k} $B > gpurun_out/r02_prof_${k}.log 2>&1
  python profiles/summarize.py gpurun_out/r02_prof_${k}.ncu-rep > gpurun_out/r02_prof_${k}.summary.csv 2>/dev/null
  grep -E "duration|dram__bytes_(read|write).sum,|issue_active" gpurun_out/r02_prof_${k}.summary.csv | tr '\n' ' '; echo
done
