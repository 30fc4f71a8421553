#!/bin/bash
# round-2 second GPU pass: device-resident iteration refactor -- parity tests, Hessian diagnostic on C3, bench with in-run parity
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=6 --durations=8 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -60 gpurun_out/r02b_pytest.log
python profiles/diag_hessian.py c3 > gpurun_out/r02b_diag_c3.log 2>&1; tail -30 gpurun_out/r02b_diag_c3.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; tail -c 2500 gpurun_out/r02b_bench_c5.json; tail -5 gpurun_out/r02b_bench_c5.err
python bench.py --steps 20 --warmup 3 --scene pile --no-cpu-baseline > gpurun_out/r02b_bench_pile.json 2> gpurun_out/r02b_bench_pile.err; tail -c 1200 gpurun_out/r02b_bench_pile.json; tail -3 gpurun_out/r02b_bench_pile.err
