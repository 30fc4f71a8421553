#!/bin/bash
# cell-centric PT kernel (vertex entries in the sorted grid): regression of every GPU test + bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02m_pytest.log 2>&1
tail -6 gpurun_out/r02m_pytest.log
run() { name=$1; shift
  env "$@" > gpurun_out/r02m_bench_$name.json 2> gpurun_out/r02m_bench_$name.err
  tail -c 300 gpurun_out/r02m_bench_$name.err | grep -v PARITY
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02m_bench_$name.json").read().strip().splitlines()[-1])
print("$name value", round(d["value"],4), "eager_profiled", round(d["config"]["eager_profiled_ms_per_step"],4), "e2e", round(d["e2e"]["value"],4), "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
}
run cellpt timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
