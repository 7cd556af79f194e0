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
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
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
act = [r for r in w[:74] if r[3]]
tot = sum(r[4] for r in act)
print("wgrad issuer totals over %d pairs: wait_full %.1f%%  wait_drained %.1f%%  issue %.1f%%  (of %d Mclk), tiles/pair min %d max %d, clk/pair min %d k max %d k" % (
    len(act), 100.0 * sum(r[0] for r in act) / tot, 100.0 * sum(r[1] for r in act) / tot, 100.0 * sum(r[2] for r in act) / tot,
    tot // 1000000, min(r[3] for r in act), max(r[3] for r in act), min(r[4] for r in act) // 1000, max(r[4] for r in act) // 1000))
t = buf.cpu().tolist()
st = [x for x in t[:256] if x]
d = [st[i + 1] - st[i] for i in range(len(st) - 1)]
print("stamps", len(st))
print("deltas", d[:80])
print("issuer: wait_epi %d wait_full %d total %d iters %d steps %d" % (t[256], t[257], t[259], t[260], t[261]))

# merged timeline of the third tile pair (see gpu_phase_clocks.py)
L.spnerf_debug_step_table.restype = ctypes.c_int
L.spnerf_debug_step_table.argtypes = [ctypes.POINTER(_cabi.NetConfig), ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
tab = (ctypes.c_int32 * (8 * 384))()
n_steps = L.spnerf_debug_step_table(ctypes.byref(model.engine.cfg), 1, tab, 384)
steps = [list(tab[8 * i:8 * i + 8]) for i in range(n_steps)]
lo, hi = [int(x) for x in os.environ.get("PC_TIMELINE", "96,132").split(",")]
ev = []
for i in range(min(n_steps, 256)):
    s_ = steps[i]
    tag = "s%03d L%d n%3d c%3d a%3d k%d%s%s" % (i, s_[6], s_[0], s_[1], s_[2], s_[3], " F" if s_[4] else "", " LAST" if s_[5] else "")
    for off, name in ((768, "wait"), (1024, "full"), (1280, "commit")):
        if t[off + i]:
            ev.append((t[off + i], i, name, tag))
t_lo = min(e[0] for e in ev if e[1] == lo)
t_hi = max(e[0] for e in ev if e[1] == min(hi, n_steps - 1))
for k, x in enumerate(t[:256]):
    if x and t_lo - 3000 <= x <= t_hi + 3000:
        ev.append((x, -1, "EPI stamp %d" % k, ""))
for e in sorted(e for e in ev if t_lo - 3000 <= e[0] <= t_hi + 3000):
    print("   %8d  %-12s %s" % (e[0] - t_lo, e[2], e[3]))
