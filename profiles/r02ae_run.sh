#!/bin/bash
# after the obstacle hand-off (filters in the classify / CCD kernels changed): full GPU regression + bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02ae_pytest.log 2>&1; tail -5 gpurun_out/r02ae_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err || tail -5 gpurun_out/r02ae_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02ae_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],4), "eager_prof", round(d["config"]["eager_profiled_ms_per_step"],4), "e2e", round(d["e2e"]["value"],4), "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
