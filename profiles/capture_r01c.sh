B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
for k in k_pairs_ee k_ccd_pairs_ee; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 2 -c 1 -f -o gpurun_out/r01c_prof_${k} $B > gpurun_out/r01c_prof_${k}.log 2>&1
done
