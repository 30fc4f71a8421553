#!/bin/bash
# round-2 ncu evidence: launch list of the bench command (eager: every launch a plain kernel launch) + one --set full capture per top kernel
# (numbers printed under ncu are not bench values)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --eager"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launch_bench.log 2>&1
python profiles/launches_by_kernel.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_by_kernel.csv
head -30 gpurun_out/r02_launches_by_kernel.csv
# -s 5: skip the launches of the warm-up (the full-CCD instance of the Tight-Inclusion passes is every second launch: odd skip count)
for k in k_elastic_grad_hess k_assemble_csr k_cell_pairs_ee k_cell_pairs_pt k_ti_stage15 k_ti_stage2 k_barrier_hessian_project k_classify_ee; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 5 -c 1 -f -o gpurun_out/r02_prof_${k} $B > gpurun_out/r02_prof_${k}.log 2>&1
  python profiles/summarize.py gpurun_out/r02_prof_${k}.ncu-rep > gpurun_out/r02_prof_${k}.summary.csv 2>/dev/null
  grep -E "duration|dram__bytes_(read|write).sum,|fp64.avg|issue_active" gpurun_out/r02_prof_${k}.summary.csv | tr '\n' ' '; echo
done
ls gpurun_out | grep -c r02_prof
