"""Timing experiments on the forward kernel (debug toggles; results of toggled runs are invalid)."""
import sys, os, types, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spnerf_b200
from spnerf_b200 import synthetic
from spnerf_b200.models import load_model
from spnerf_b200 import engine as E, config
cfg = config.make_args(sem=True, num_sem_classes=3)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = load_model(types.SimpleNamespace(**vars(cfg))).to(dev)
B, N = 8192, 64
batch = synthetic.make_batch(B, seed=5, device=dev)
rays = batch["rays"]
z = E.sample_coarse(rays, torch.rand(B, N, device=dev), N)
eng = model.engine
res = {}
for flags in (0, 1, 2, 3, 4, 5, 6, 7):
    for _ in range(2):
        eng.forward(rays, N, z=z, labels=batch["sems"], debug_flags=flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(rays, N, z=z, labels=batch["sems"], debug_flags=flags)
    e1.record(); torch.cuda.synchronize()
    res[flags] = e0.elapsed_time(e1) / 5
    print("flags", flags, "noload" if flags & 1 else "", "noepi" if flags & 2 else "", "nomma" if flags & 4 else "",
          "ms", round(res[flags], 3), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fwd_timing.json"), "w"))
