#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "c3 or c4" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
# per-launch durations of the same bench command (numbers printed under ncu are NOT bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02c_launch_bench.log 2>&1
tail -2 gpurun_out/r02c_launch_bench.log | cut -c1-300
