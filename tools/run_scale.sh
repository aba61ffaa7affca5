# usage: bash tools/run_scale.sh N   (on an N-GPU box under gpurun --gpus N): multi-GPU parity test, then the bench line
N=$1
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
fi
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_c4_n$N.json") if l.startswith("{")][-1])
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "exact", d.get("exact_gram",{}).get("value"))
print(d["phases_ms_last_iter"])
PY
tail -2 gpurun_out/bench_c4_n$N.err
