# usage: bash tools/run_scale2.sh TAG N1 [N2 ...]   (under gpurun --gpus max(N)): multi-GPU parity check at the box's GPU count,
# then one bench line per N (BSFG_OVERLAP as set by the caller); lines land in gpurun_out/TAG_bench_c4_n<N>.json
TAG=$1; shift
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/${TAG}_multi_gpu_test.log 2>&1; tail -3 gpurun_out/${TAG}_multi_gpu_test.log
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-exact-leg > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-exact-leg > gpurun_out/${TAG}_bench_c4_n$N.json 2> gpurun_out/${TAG}_bench_c4_n$N.err
  fi
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench_c4_n$N.json") if l.startswith("{")][-1])
print("N=$N value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d["e2e"].get("ms_per_step_parts"))
print({k: round(v,3) for k,v in d["phases_ms_last_iter"].items()})
PY
  tail -2 gpurun_out/${TAG}_bench_c4_n$N.err
done
