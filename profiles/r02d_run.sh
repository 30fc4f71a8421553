#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/diag_tet_hessian.py > gpurun_out/r02d_diag_tet.log 2>&1; tail -20 gpurun_out/r02d_diag_tet.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --durations=5 > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02d_pytest.log | tail -15
for mode in "1 1" "0 1" "1 0"; do set -- $mode
  IPCGPU_TI_MODE=$1 IPCGPU_TET_KERNEL=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_ti$1_tet$2.json 2> gpurun_out/r02d_bench_ti$1_tet$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02d_bench_ti$1_tet$2.json").read().strip().splitlines()[-1])
print("TI_MODE=$1 TET=$2", round(d["value"],3), round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["parity"]["ok"] if d["parity"] else None)
PY
done
