set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; tail -c 300 gpurun_out/bench_r2_d.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2_d.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels_ms"], d["roofline"]["frac"], d["roofline"]["kernel"])
print({k:(v["ms_per_step"] if "ms_per_step" in v else v["seconds_per_image"]) for k,v in d["other_configs"].items()})
PY
python tools/gpu_determinism.py 2>&1 | tail -5
