# usage (under gpurun, one GPU): bash tools/profile_iteration.sh TAG
# One warm Gibbs iteration of `bench.py --steps 2 --warmup 3` (C4): (1) every launch with its device time,
# (2) an `ncu --set full` capture of every kernel of that iteration, exported on the box as the raw-metrics CSV (the report
# itself is ~100 MB, over the 64 MiB that travel back; it is kept only when small).  The iteration is bracketed inside the
# library (BSFG_PROFILE_ITER, cudaProfilerStart/Stop).
TAG=$1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact-leg"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
BSFG_PROFILE_ITER=4 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
[ "${PROFILE_FULL:-1}" = "1" ] || exit 0
BSFG_PROFILE_ITER=4 ncu --profile-from-start off --set full --clock-control none \
    -o /tmp/prof_${TAG}_iteration $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/prof_${TAG}_iteration.ncu-rep --page raw --csv > gpurun_out/${TAG}_iteration_raw.csv 2>/dev/null
ls -la /tmp/prof_${TAG}_iteration.ncu-rep gpurun_out/${TAG}_iteration_raw.csv gpurun_out/${TAG}_launches.csv
SZ=$(stat -c %s /tmp/prof_${TAG}_iteration.ncu-rep)
if [ "$SZ" -lt 45000000 ]; then cp /tmp/prof_${TAG}_iteration.ncu-rep gpurun_out/; fi
du -sh gpurun_out
