#!/bin/bash
# re-entry baseline: GPU tests, bench line (both arms), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02g_smi.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02g_pytest.log 2>&1
tail -5 gpurun_out/r02g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -c 600 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
print("roof", d["roofline"]["frac"], d["roofline_hessian_to_csr"]["frac"], d["roofline_ccd_narrow"]["frac"], d.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02g_launch_bench.log 2>&1
python profiles/launches_by_kernel.py gpurun_out/r02g_launches.csv > gpurun_out/r02g_launches_by_kernel.csv 2>gpurun_out/r02g_sum.err; head -40 gpurun_out/r02g_launches_by_kernel.csv
