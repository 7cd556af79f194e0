set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r2_4gpu.json 2> gpurun_out/bench_r2_4gpu.err; tail -c 300 gpurun_out/bench_r2_4gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --skip-extra > gpurun_out/bench_r2_1gpu_same_box.json 2>/dev/null
python - <<'PY'
import json
def last(p):
    for line in open(p):
        if line.startswith("{"): d=json.loads(line)
    return d
a=last("gpurun_out/bench_r2_4gpu.json"); b=last("gpurun_out/bench_r2_1gpu_same_box.json")
print("4gpu", a["value"], a["ms_per_step"], "e2e", a["e2e"]["value"], a["e2e"]["ms_per_step"], a["kernels_ms"].get("allreduce"))
print("1gpu", b["value"], b["ms_per_step"], "e2e", b["e2e"]["value"])
print("eff", a["value"]/(4*b["value"]), "e2e eff", a["e2e"]["value"]/(4*b["e2e"]["value"]))
PY
