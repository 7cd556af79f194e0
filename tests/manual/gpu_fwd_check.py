"""GPU check of the fused point-network forward against the oracle network evaluated with torch
fp32 on the same device, plus a first timing.  usage: python tools/gpu_fwd_check.py [case ...]"""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {
    "sem": dict(sem=True, num_sem_classes=3, mapping=False, beta=False),
    "sem_map": dict(sem=True, num_sem_classes=3, mapping=True, beta=False),
    "plain": dict(sem=False, mapping=False, beta=False),
    "beta_map": dict(sem=True, num_sem_classes=3, mapping=True, beta=True),
    "time_sem": dict(sem=True, num_sem_classes=3, mapping=False, beta=False, time=True),
}


def run(name):
    import torch
    import spnerf_b200
    from spnerf_b200 import synthetic
    from spnerf_b200.models import load_model
    from oracle import spnerf_oracle as O

    spec = dict(CASES[name])
    timing = spec.pop("time", False)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.make_cfg(**spec)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = load_model(types.SimpleNamespace(**vars(cfg)))
    with torch.no_grad():
        model.sigma_from_xyz[0].bias.fill_(3.0)
        model.sigma_from_xyz[0].weight.mul_(4.0)
    model = model.to(dev)
    B, N = (8192, 64) if timing else (300, 64)     # 300*64 = 19200 points: 150 tiles -> 2 tiles on some CTAs
    batch = synthetic.make_batch(B, seed=5, device=dev)
    rays = batch["rays"]
    g = torch.Generator().manual_seed(1)
    z = O.stratified_z(rays.cpu(), N, torch.rand(B, N, generator=g)).to(dev).contiguous()
    labels = batch["sems"] if cfg.sem else None
    t_emb = torch.randn(B, cfg.t_embbeding_tau, device=dev) if cfg.beta else None
    eng = model.engine
    out, _ = eng.forward(rays, N, z=z, labels=labels, t_emb=t_emb, save=False)
    torch.cuda.synchronize()
    res = {"case": name, "watchdog": int(spnerf_b200._cabi.lib().spnerf_watchdog_code())}
    # oracle on the same device
    P = {k: v.detach() for k, v in model.named_parameters()}
    xyz = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3)
    with torch.no_grad():
        want = O.point_network(P, cfg, xyz, torch.repeat_interleave(rays[:, 8:11], N, 0),
                               None if labels is None else torch.repeat_interleave(labels, N, 0),
                               None if t_emb is None else torch.repeat_interleave(t_emb, N, 0))
    err = (out - want).abs()
    res["max_abs_err_per_col"] = [float(x) for x in err.max(0).values]
    res["nan"] = int(torch.isnan(out).sum())
    res["ok"] = bool(err.max() < 5e-3 and res["nan"] == 0)
    # training mode must give the same rows and fill the save area
    out2, saves = eng.forward(rays, N, z=z, labels=labels, t_emb=t_emb, save=True)
    torch.cuda.synchronize()
    res["save_mode_equal"] = bool(torch.equal(out, out2))
    res["save_bytes"] = int(saves.numel())
    if timing:
        for save in (False, True):
            for _ in range(3):
                eng.forward(rays, N, z=z, labels=labels, t_emb=t_emb, save=save)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                eng.forward(rays, N, z=z, labels=labels, t_emb=t_emb, save=save)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            flops = 2 * 2632192 * B * N
            res[f"ms_save{int(save)}"] = ms
            res[f"tflops_save{int(save)}"] = flops / ms / 1e9
    print(json.dumps(res))
    return 0 if res["ok"] else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        sys.exit(run(sys.argv[1]))
    allres = {}
    for name in CASES:
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
            allres[name] = json.loads(lines[-1]) if lines else {"ok": False, "rc": p.returncode,
                                                                 "stderr": p.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            allres[name] = {"ok": False, "timeout": True}
        print(name, json.dumps(allres[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fwd_check.json"), "w") as f:
        json.dump(allres, f, indent=1)


if __name__ == "__main__":
    main()
