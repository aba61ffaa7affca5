# launch list of OUR kernels inside bench.py (the torch setup kernels are filtered out by name)
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
RX='regex:^(dgemm_tn|expsum_|splitk|transpose_kernel|scale_kernel|gram_features|lambda_|chol_|pack_|factor_|means_|h2_|grid_draw|prec_|fa_|tau_e|delta_|lt_to|plam_|trait_|colsumsq|diag_res|build_prior|addk|compact)'
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | head -c 300
