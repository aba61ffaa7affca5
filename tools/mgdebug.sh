for m in 0 1 2; do
echo "=== BSFG_SHARD_REPLICATED=$m"
BSFG_SHARD_REPLICATED=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$m tests/multi_gpu_check.py 2>&1 | grep "rank0\]\|multi_gpu_check ok" | grep -v "site-packages\|Warning" | tail -6
done
