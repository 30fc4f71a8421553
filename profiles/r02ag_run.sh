#!/bin/bash
# regression at N = $1 ranks after the last changes of the round (pair-kernel staging, obstacle pair rules); N = 1 as in r02y
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02ag_pytest.log 2>&1; tail -3 gpurun_out/r02ag_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ag_bench_1gpu.json 2> gpurun_out/r02ag_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ag_bench_reference_arm.json 2> gpurun_out/r02ag_bench_reference_arm.err
python -c "
import json
d=json.loads(open('gpurun_out/r02ag_bench_reference_arm.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02ag_bench_${N}gpu.json 2> gpurun_out/r02ag_bench_${N}gpu.err
fi
grep PARITY gpurun_out/r02ag_bench_${N}gpu.err | tail -1 | cut -c1-160
python - <<PY
import json
d=json.loads(open("gpurun_out/r02ag_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"],4), "eager_prof", round(d["config"]["eager_profiled_ms_per_step"],4), "e2e", round(d["e2e"]["value"],4), "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
print("roof", round(d["roofline"]["frac"],3), round(d["roofline_hessian_to_csr"]["frac"],3), round(d["roofline_ccd_narrow"]["frac"],4), d.get("cpu_baseline") and d["cpu_baseline"]["value"], d["clocks"])
PY
