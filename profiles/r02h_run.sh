#!/bin/bash
# N-rank bench line (N = $1) with in-run parity and the per-stage table
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02h_bench_${N}gpu.json 2> gpurun_out/r02h_bench_${N}gpu.err
grep PARITY gpurun_out/r02h_bench_${N}gpu.err | tail -1
tail -c 400 gpurun_out/r02h_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02h_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d["parity"] and d["parity"]["ok"], "sum_stage", d["sum_stage_ms"])
print({k: round(v,4) for k,v in d["stage_ms"].items()})
PY
