"""Clock log of the fused forward / backward-data kernel (CTA 0): cycles between consecutive epilogue
stamps, the MMA issuers' accumulated wait times, and a merged per-step timeline of the third tile pair.
usage: gpu_phase_clocks.py [debug_flags[s]...]   ('s' = training mode, activations saved)
env: PC_RAYS, PC_TIMELINE=first,last step of the timeline print (default 40,80)"""
import sys, os, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, _cabi
from spnerf_b200 import engine as E, config

dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
B, N = int(os.environ.get('PC_RAYS', '8192')), 64
batch = synthetic.make_batch(B, seed=269, device=dev)
z = E.sample_coarse(batch["rays"], torch.rand(B, N, device=dev), N)
eng = model.engine
L = _cabi.lib()
L.spnerf_debug_phase_clocks_fwd.restype = None
L.spnerf_debug_phase_clocks_fwd.argtypes = [ctypes.c_void_p]
L.spnerf_debug_step_table.restype = ctypes.c_int
L.spnerf_debug_step_table.argtypes = [ctypes.POINTER(_cabi.NetConfig), ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
tab = (ctypes.c_int32 * (8 * 384))()
n_steps = L.spnerf_debug_step_table(ctypes.byref(eng.cfg), 0, tab, 384)
steps = [list(tab[8 * i:8 * i + 8]) for i in range(n_steps)]
lo, hi = [int(x) for x in os.environ.get("PC_TIMELINE", "40,80").split(",")]
out = {}
for spec in (sys.argv[1:] or ["0s", "0"]):
    save = spec.endswith("s")
    flags = int(spec.rstrip("s"))
    for _ in range(2):
        eng.forward(batch["rays"], N, z=z, labels=batch["sems"], save=save, debug_flags=flags)
    torch.cuda.synchronize()
    buf = torch.zeros(2048, dtype=torch.int64, device=dev)
    L.spnerf_debug_phase_clocks_fwd(buf.data_ptr())
    eng.forward(batch["rays"], N, z=z, labels=batch["sems"], save=save, debug_flags=flags)
    torch.cuda.synchronize()
    L.spnerf_debug_phase_clocks_fwd(None)
    t = buf.cpu().tolist()
    st = [x for x in t[:256] if x]
    d = [st[i + 1] - st[i] for i in range(len(st) - 1)]
    per_tile = 31
    print(f"== flags {flags} save {save}: first tile deltas", d[:per_tile])
    print("   second tile", d[per_tile:2 * per_tile])
    print("   issuer0: wait_epi %d wait_full %d total %d iters %d steps %d" % (t[256], t[257], t[259], t[260], t[261]))
    # merged timeline of the third tile pair
    ev = []
    for i in range(min(n_steps, 256)):
        s = steps[i]
        tag = "s%03d L%d n%3d c%3d a%3d k%d%s%s%s" % (i, s[6], s[0], s[1], s[2], s[3], " F" if s[4] else "", " LAST" if s[5] else "", " early" if s[7] else "")
        for off, name in ((512, "prod-empty"), (768, "gate"), (1024, "full"), (1280, "commit")):
            if t[off + i]:
                ev.append((t[off + i], i, name, tag))
    t_lo = min(e[0] for e in ev if e[1] == lo)
    t_hi = max(e[0] for e in ev if e[1] == min(hi, n_steps - 1))
    for k, x in enumerate(t[:256]):
        if x and t_lo - 3000 <= x <= t_hi + 3000:
            ev.append((x, -1, "EPI stamp %d" % k, ""))
    for k, x in enumerate(t[1536:1792]):
        if x and t_lo - 3000 <= x <= t_hi + 3000:
            ev.append((x, -1, "EPI half %d" % k, ""))
    ev = sorted(e for e in ev if t_lo - 3000 <= e[0] <= t_hi + 3000 and e[2] != "prod-empty")
    for e in ev:
        print("   %8d  %-12s %s" % (e[0] - t_lo, e[2], e[3]))
    out[spec] = {"deltas": d, "issuer": t[256:262]}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "phase_clocks.json"), "w"))
