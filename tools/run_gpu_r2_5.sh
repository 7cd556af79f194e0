set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 3000 gpurun_out/bench_r2_a.err; cat gpurun_out/bench_r2_a.json | cut -c1-6000
