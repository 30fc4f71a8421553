#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one --set full capture per top kernel (numbers printed under ncu are not bench values)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launch_bench.log 2>&1
for k in k_elastic_grad_hess_rolled k_assemble_csr k_ti_groups k_ti_stage2 k_ccd_pairs_ee k_pairs_ee; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 3 -c 1 -f -o gpurun_out/r02_prof_${k} $B > gpurun_out/r02_prof_${k}.log 2>&1
  python profiles/summarize.py gpurun_out/r02_prof_${k}.ncu-rep > gpurun_out/r02_prof_${k}.summary.csv 2>/dev/null
done
ls -la gpurun_out/ | grep r02_
