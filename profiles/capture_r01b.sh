set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01b_launches.csv $B > gpurun_out/r01b_launch_bench.log 2>&1
for k in k_elastic_grad_hess k_ti_stage1 k_barrier_hessian_project k_pairs_ee k_ccd_pairs_ee k_ti_stage2; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 2 -c 1 -f -o gpurun_out/r01b_prof_${k} $B > gpurun_out/r01b_prof_${k}.log 2>&1
done
ls -la gpurun_out/
