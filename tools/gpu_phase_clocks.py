"""Per-phase clock log of the fused forward / backward-data kernels (block 0, first tiles):
prints cycles between consecutive epilogue stamps (begin = MMA phase retired, end = epilogue done)."""
import sys, os, types, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, train_step, _cabi

dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batch = synthetic.make_batch(rays, seed=269, device=dev)
L = _cabi.lib()
for fn in (L.spnerf_debug_phase_clocks_fwd, L.spnerf_debug_phase_clocks_bwd):
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p]
for _ in range(2):
    train_step.fused_step(model, args, batch, repack=True)
torch.cuda.synchronize()
bf = torch.zeros(256, dtype=torch.int64, device=dev)
bb = torch.zeros(256, dtype=torch.int64, device=dev)
L.spnerf_debug_phase_clocks_fwd(bf.data_ptr())
L.spnerf_debug_phase_clocks_bwd(bb.data_ptr())
train_step.fused_step(model, args, batch, repack=True)
torch.cuda.synchronize()
L.spnerf_debug_phase_clocks_fwd(None)
L.spnerf_debug_phase_clocks_bwd(None)
out = {}
for name, b in (("fwd", bf), ("bwd", bb)):
    t = b.cpu().tolist()
    t = [x for x in t if x]
    d = [t[i + 1] - t[i] for i in range(len(t) - 1)]
    out[name] = d
    print(name, "stamps", len(t), "total", t[-1] - t[0] if t else 0)
    print(" deltas:", d[:120])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "phase_clocks.json"), "w"))
