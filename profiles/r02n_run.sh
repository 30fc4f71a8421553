#!/bin/bash
# launch list of the current bench command (eager so that every launch is a plain kernel launch for ncu)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --eager > gpurun_out/r02n_launch_bench.log 2>&1
python profiles/launches_by_kernel.py gpurun_out/r02n_launches.csv > gpurun_out/r02n_launches_by_kernel.csv; head -45 gpurun_out/r02n_launches_by_kernel.csv
