# what the driver runs at round end, on one box: GPU tests, smoke, bench (both arms)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'], d['gpu_launches'], d['steps'], d['warmup'])"
timeout 900 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_default_ref.json
