#!/bin/bash
# friction/inertia GPU tests, regression of the contact + CCD + full-size tests after the pipelined pair kernels and the regrouped CSR assembly, bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_friction.py -x -q ) > gpurun_out/r02i_pytest_friction.log 2>&1
tail -15 gpurun_out/r02i_pytest_friction.log
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_friction.py ) > gpurun_out/r02i_pytest.log 2>&1
tail -5 gpurun_out/r02i_pytest.log
for u in 4 1; do
IPCGPU_ASM_UNROLL=$u timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_u$u.json 2> gpurun_out/r02i_bench_u$u.err
tail -c 300 gpurun_out/r02i_bench_u$u.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02i_bench_u$u.json").read().strip().splitlines()[-1])
print("unroll $u value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
done
