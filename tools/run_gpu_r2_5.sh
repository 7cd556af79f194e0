set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 --skip-extra > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -c 500 gpurun_out/bench_r2_b.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2_b.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels_ms"], d["roofline"]["frac"], d["small_batch"])
PY
