# launch list of one bench run, then ONE full ncu capture of the exponential-sum trait product (the dgemm_tn launch that
# follows the node product, i.e. the second dgemm_tn after an expsum_traits_kernel), located from the launch list itself.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact-leg"
RX='regex:^(dgemm_tn|expsum_|splitk|transpose_kernel|scale_kernel|gram_features|lambda_|chol_|pack_|factor_|means_|h2_|grid_draw|prec_|fa_|tau_e|delta_|lt_to|plam_|trait_|colsumsq|diag_res|build_prior|addk|compact)'
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
IDX=$(python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ki = rows[hi].index("Kernel Name")
names = [r[ki] for r in rows[hi + 1:]]
last = [i for i, nm in enumerate(names) if nm.startswith("expsum_traits")][-1]
dg = [i for i, nm in enumerate(names) if "dgemm_tn_kernel" in nm]
after = [i for i in dg if i > last]
print(dg.index(after[1]))          # 0-based index among dgemm_tn launches: node product, then the trait product
PY
)
echo "trait-product dgemm index: $IDX" | tee gpurun_out/qgemm_index.txt
ncu --set full --clock-control none --import-source on -k regex:dgemm_tn_kernel -s $IDX -c 1 -f -o gpurun_out/prof_r01b_qgemm $CMD > gpurun_out/ncu_full_qgemm.log 2>&1
tail -3 gpurun_out/ncu_full_qgemm.log
