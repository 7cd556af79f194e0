"""Bitwise run-to-run determinism of the fused forward / full training step (same inputs, same draws)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, train_step
from spnerf_b200 import engine as E, config

dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
B, N = 8192, 64
batch = synthetic.make_batch(B, seed=269, device=dev)
z = E.sample_coarse(batch["rays"], torch.rand(B, N, device=dev), N)
eng = model.engine
ref = None
for i in range(6):
    out, _ = eng.forward(batch["rays"], N, z=z, labels=batch["sems"], save=(i % 2 == 1))
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    else:
        d = (out - ref).abs().max().item()
        print("fwd run", i, "save", i % 2 == 1, "bit-equal", bool(torch.equal(out, ref)), "max abs diff", d)
gref = None
for i in range(4):
    torch.manual_seed(3)
    E.manual_seed(3)
    flat, _, scalars, _ = train_step.fused_step(model, args, batch, repack=True)
    torch.cuda.synchronize()
    if gref is None:
        gref = flat.clone()
    else:
        print("step run", i, "grad bit-equal", bool(torch.equal(flat, gref)), "rel diff",
              ((flat - gref).norm() / gref.norm()).item())
