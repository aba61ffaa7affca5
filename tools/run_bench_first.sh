set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload tiny --steps 5 --warmup 3 2>&1 | tail -3
python bench.py --workload C2 --steps 10 --warmup 3 --cpu-sample-traits 200 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4_first.json 2> gpurun_out/bench_c4_first.err; tail -c 3000 gpurun_out/bench_c4_first.json; tail -5 gpurun_out/bench_c4_first.err
