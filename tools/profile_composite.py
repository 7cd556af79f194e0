"""ncu driver: compositing forward / adjoint on a batch that leaves L2 (262144 rays x 64 samples)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spnerf_b200
from spnerf_b200 import engine as E
dev = torch.device("cuda:0")
nr, N = 262144, 64
out = torch.rand(nr * N, 11, device=dev); out[:, 3] *= 30
z = torch.sort(torch.rand(nr, N, device=dev) * 0.2, -1).values.contiguous()
g_rgb, g_depth, g_sem = torch.randn(nr, 3, device=dev), torch.randn(nr, device=dev), torch.randn(nr, 3, device=dev)
for _ in range(2):
    w, t_, rgb, raw, depth, sem = E.composite_fwd(out, z, 11, 8, 3)
    E.composite_bwd(out, z, w, t_, raw, 11, 8, 3, g_rgb=g_rgb, g_depth=g_depth, g_sem=g_sem)
torch.cuda.synchronize()
print("done")
