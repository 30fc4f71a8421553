#!/bin/bash
# slot-major Hessian intermediate: regression + A/B bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02t_pytest.log 2>&1; tail -4 gpurun_out/r02t_pytest.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_$name.json 2> gpurun_out/r02t_bench_$name.err
  tail -c 200 gpurun_out/r02t_bench_$name.err | grep -v PARITY
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t_bench_$name.json").read().strip().splitlines()[-1])
s=d["stage_ms"]
print("$name value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "tet", round(s["elastic_tet"],4), "asm", round(s["assemble_csr"],4), "gather", round(s["gather_gradient"],4), "parity", d["parity"] and d["parity"]["ok"], "h2csr frac", round(d["roofline_hessian_to_csr"]["frac"],3))
PY
}
run slot IPCGPU_HESS_LAYOUT=1
run tile IPCGPU_HESS_LAYOUT=0
