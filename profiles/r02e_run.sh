#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --durations=5 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02e_pytest.log | tail -15
for mode in 2 0; do
  IPCGPU_TI_MODE=$mode timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_ti$mode.json 2> gpurun_out/r02e_bench_ti$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench_ti$mode.json").read().strip().splitlines()[-1])
print("TI_MODE=$mode", round(d["value"],3), round(d["e2e"]["value"],3), d["gpu_launches"]/20, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["parity"]["ok"] if d["parity"] else None)
print(d["config"]["full_ccd_candidates_survivors_warnings_deferred_boxesThreadPass_boxesWarpPass_longestPairCycles_totalCycles"])
PY
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --scene pile > gpurun_out/r02e_bench_pile.json 2> gpurun_out/r02e_bench_pile.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench_pile.json").read().strip().splitlines()[-1])
print("pile", round(d["value"],3), round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["parity"]["ok"] if d["parity"] else None)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02e_launch_bench.log 2>&1
