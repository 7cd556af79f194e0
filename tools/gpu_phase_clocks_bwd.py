"""Per-phase clock log of the fused backward-data kernel (block 0)."""
import sys, os, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, train_step, _cabi
dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
batch = synthetic.make_batch(8192, seed=269, device=dev)
L = _cabi.lib()
L.spnerf_debug_phase_clocks_bwd.restype = None
L.spnerf_debug_phase_clocks_bwd.argtypes = [ctypes.c_void_p]
for _ in range(2):
    train_step.fused_step(model, args, batch, repack=True)
torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
wb = torch.zeros(8 * 80, dtype=torch.int64, device=dev)
L.spnerf_debug_counters_wgrad.restype = None
L.spnerf_debug_counters_wgrad.argtypes = [ctypes.c_void_p]
L.spnerf_debug_counters_wgrad(wb.data_ptr())
L.spnerf_debug_phase_clocks_bwd(buf.data_ptr())
train_step.fused_step(model, args, batch, repack=True)
torch.cuda.synchronize()
L.spnerf_debug_phase_clocks_bwd(None)
L.spnerf_debug_counters_wgrad(None)
w = wb.cpu().view(80, 8).tolist()
print("wgrad pairs: (tiles, total kclk, clk/tile)", [(r[3], r[4] // 1000, r[4] // max(r[3], 1)) for r in w[:74]])
t = buf.cpu().tolist()
st = [x for x in t[:256] if x]
d = [st[i + 1] - st[i] for i in range(len(st) - 1)]
print("stamps", len(st))
print("deltas", d[:80])
print("issuer: wait_epi %d wait_full %d total %d iters %d steps %d" % (t[256], t[257], t[259], t[260], t[261]))
