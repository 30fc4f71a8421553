#!/bin/bash
# round-2 first GPU pass: parity tests (incl. the new full-size configs) + bench lines of both scenes
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpu.txt; nproc >> gpurun_out/r02a_gpu.txt
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -30 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err; tail -c 3000 gpurun_out/r02a_bench_c5.json
python bench.py --steps 20 --warmup 3 --scene pile --no-cpu-baseline > gpurun_out/r02a_bench_pile.json 2> gpurun_out/r02a_bench_pile.err; tail -c 1500 gpurun_out/r02a_bench_pile.json
