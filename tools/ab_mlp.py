"""A/B timing of the three point-network kernels inside the fused training step (BASELINE config 2 shape).
Usage: python tools/ab_mlp.py [tag] [stagger settings ...]   e.g.  python tools/ab_mlp.py base 0,1 24000,2 24000,8
Each setting is the SPNERF_STAGGER value of an SPNERF_EXPERIMENTS build; the library is picked by SPNERF_LIB."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200  # noqa: F401
from spnerf_b200 import synthetic, train_step, engine as E

tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
settings = sys.argv[2:] or ["0,1"]
rays = int(os.environ.get("AB_RAYS", "8192"))
reps = int(os.environ.get("AB_REPS", "8"))
dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
batch = synthetic.make_batch(rays, seed=269, device=dev)
out = {}
for rnd in range(2):                      # two interleaved rounds: box drift shows up as a difference between them
    for st in settings:
        stg, _, dbg = st.partition(";")          # "stagger[;debug flags]"
        os.environ["SPNERF_STAGGER"] = stg
        E._ENV_DEBUG = int(dbg or "0")
        for _ in range(2):
            train_step.fused_step(model, args, batch, repack=True)
        torch.cuda.synchronize()
        acc = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(reps):
            t = train_step.StepTimer(True)
            train_step.fused_step(model, args, batch, repack=True, timer=t)
            torch.cuda.synchronize()
            for k, v in t.durations_ms().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        e0.record()
        for _ in range(reps):
            train_step.fused_step(model, args, batch, repack=True)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / reps
        row = {k: round(acc[k], 3) for k in ("mlp_fwd", "mlp_bwd_data", "mlp_bwd_weights")}
        row["step"] = round(step_ms, 3)
        out[f"{tag}|{st}|r{rnd}"] = row
        print(tag, st, "round", rnd, row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"ab_{tag}.json"), "w") as f:
    json.dump(out, f, indent=1)
