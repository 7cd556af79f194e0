"""Per-phase clock log of the fused forward kernel (block 0): cycles between consecutive epilogue
stamps, plus the MMA issuer's accumulated wait times.  usage: gpu_phase_clocks.py [debug_flags...]"""
import sys, os, types, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200
from spnerf_b200 import synthetic, _cabi
from oracle import spnerf_oracle as O

dev = torch.device("cuda:0")
args = bench.make_args()
model = bench.build_model(args, dev)
B, N = int(os.environ.get('PC_RAYS', '8192')), 64
batch = synthetic.make_batch(B, seed=269, device=dev)
z = O.stratified_z(batch["rays"], N, torch.rand(B, N, device=dev)).contiguous()
eng = model.engine
L = _cabi.lib()
L.spnerf_debug_phase_clocks_fwd.restype = None
L.spnerf_debug_phase_clocks_fwd.argtypes = [ctypes.c_void_p]
out = {}
for spec in (sys.argv[1:] or ["0s", "0", "7", "5", "1"]):
    save = spec.endswith("s")
    flags = int(spec.rstrip("s"))
    for _ in range(2):
        eng.forward(batch["rays"], N, z=z, labels=batch["sems"], save=save, debug_flags=flags)
    torch.cuda.synchronize()
    buf = torch.zeros(512, dtype=torch.int64, device=dev)
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
    print("   issuer: wait_epi %d wait_full %d wait_pfull %d total %d iters %d steps %d | producer wait_empty %d" % (
        t[256], t[257], t[258], t[259], t[260], t[261], t[264]))
    base = min(x for x in t[300:452] if x)
    rel = lambda a: [x - base if x else -1 for x in a]
    print("   steps 40..71: producer empty-seen ", rel(t[300:332]))
    print("                 issuer full-seen    ", rel(t[340:372]))
    print("                 issuer commit       ", rel(t[420:452]))
    out[spec] = {"deltas": d, "issuer": t[256:262], "producer_wait_empty": t[264]}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "phase_clocks.json"), "w"))
