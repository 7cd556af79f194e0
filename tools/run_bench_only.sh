cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2_e.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels_ms"], d["roofline"]["frac"], d["clocks"])
print({k:(v["ms_per_step"] if "ms_per_step" in v else v["seconds_per_image"]) for k,v in d["other_configs"].items()})
PY
