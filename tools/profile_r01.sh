# Round-1 profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture
# of the dominant kernel (the weighted-Gram dgemm_tn launch = 9th dgemm_tn_kernel launch of bench.py).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/plain.log | head -c 600
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dgemm_tn_kernel -s 8 -c 1 -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
