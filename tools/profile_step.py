"""Small driver for ncu: a few fused training steps (BASELINE config 2 shape) and nothing else."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, train_step
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
batch = synthetic.make_batch(rays, seed=269, device=dev)
for _ in range(steps):
    train_step.fused_step(model, args, batch, repack=True)
torch.cuda.synchronize()
print("profile_step done")
