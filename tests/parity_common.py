"""Shared helpers of the parity tests: load a golden case (tests/golden/*.npz, written by
oracle/make_golden.py from the unmodified reference), rebuild its model and inputs, run the CUDA
product path through the reference-shaped API and report per-quantity errors."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# stated tolerances (BASELINE.json north_star): rgb max-abs 1e-3, depth 1e-2 m (scene range 141.21875 m
# per normalised unit); fp16-operand / fp32-accumulate network
TOL = {
    "rgb": 1e-3,
    "depth": 1e-2 / 141.21875,
    "weights": 2e-3,
    "transparency": 2e-3,
    "albedo": 2e-3,
    "sun": 2e-3,
    "sky": 1e-5,
    "sem_logits": 5e-3,
    "beta": 2e-3,
}


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    return g, meta


def make_args(meta):
    cfg = dict(meta["cfg"])
    cfg["skips"] = tuple(cfg["skips"])
    return types.SimpleNamespace(**cfg)


def build_model(meta, device):
    """Same construction as oracle/make_golden.py: reference init stream under manual_seed(0),
    optional 'trained-like' density head."""
    from spnerf_b200.models import load_model
    args = make_args(meta)
    torch.manual_seed(0)
    if getattr(args, "siren", True):
        model = load_model(args)
    else:               # the ReLU variant is only reachable through the class (load_model never passes siren)
        from spnerf_b200.models import SPNeRF
        model = SPNeRF(num_sem_classes=args.num_sem_classes, s_embedding_factor=args.s_embedding_factor,
                       layers=args.fc_layers, feat=args.fc_units, mapping=args.mapping,
                       t_embedding_dims=args.t_embbeding_tau, beta=args.beta, sem=args.sem, siren=False)
    t_table = None
    if args.beta:
        t_table = torch.nn.Embedding(30, args.t_embbeding_tau)
    if meta["trained_like"]:
        with torch.no_grad():
            model.sigma_from_xyz[0].bias.fill_(3.0)
            model.sigma_from_xyz[0].weight.mul_(4.0)
    model = model.to(device)
    if t_table is not None:
        t_table = t_table.to(device)
    return model, t_table, args


def state_hash(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


class Draws:
    def __init__(self, g, device):
        self.u = [torch.from_numpy(g[f"uniform_{i}"]).to(device) for i in range(10) if f"uniform_{i}" in g]
        self.n = [torch.from_numpy(g[f"normal_{i}"]).to(device) for i in range(10) if f"normal_{i}" in g]

    def uniform(self, shape):
        t = self.u.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t

    def normal(self, shape):
        t = self.n.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t


def run_case(name, device="cuda:0", with_backward=True):
    """Returns a report dict: per-output max-abs error vs the golden file, loss errors, gradient errors."""
    from spnerf_b200.modules.rendering import render_rays
    from spnerf_b200.modules import metrics
    g, meta = load_case(name)
    model, t_emb_mod, args = build_model(meta, device)
    rep = {"case": name, "state_ok": state_hash({k: v.cpu() for k, v in model.state_dict().items()}) == meta["state_sha256"]}
    ins = {k[3:]: torch.from_numpy(g[k]).to(device) for k in g.files if k.startswith("in_")}
    args._rng = Draws(g, device)
    mode = meta["mode"]
    train = mode == "train"
    models = {"coarse": model}
    if args.beta:
        models["t"] = t_emb_mod
    kw = dict(semantics=ins["sems"] if args.sem else None, mode=mode)
    if train:
        kw.update(valid_depth=ins["valid_depth"], target_depths=ins["depths"], target_std=ins["depth_std"])
    res = render_rays(models, args, ins["rays"], ins["ts"] if args.beta else None, **kw)
    torch.cuda.synchronize()
    rep["keys_ok"] = sorted(res) == sorted(k[4:] for k in g.files if k.startswith("out_"))
    out_err = {}
    for k, v in res.items():
        want = torch.from_numpy(g["out_" + k]).to(device)
        if k.startswith("z_vals"):
            out_err[k] = {"bit_equal": bool(torch.equal(v, want)), "max_abs": float((v - want).abs().max())}
        else:
            out_err[k] = {"max_abs": float((v.detach() - want).abs().max()), "nan": int(torch.isnan(v).sum())}
    rep["out"] = out_err
    if not with_backward:
        return rep
    # losses exactly as oracle/make_golden.py
    loss_fn = metrics.load_loss(args)
    loss, ld = loss_fn(res, ins["rgbs"])
    if train:
        dl = metrics.DepthLoss(lambda_ds=1.0, GNLL=False, usealldepth=False)
        l2, d2 = dl(res, ins["depths"][:, 0], ins["depths"][:, 1], target_valid_depth=ins["valid_depth"],
                    target_std=ins["depth_std"])
        loss = loss + l2
        ld.update(d2)
    if args.sem:
        sl = metrics.SemanticLoss(lambda_ss=1.0)
        l3, d3 = sl(res, ins["sems"])
        loss = loss + l3
        ld.update(d3)
    rep["loss"] = {k: {"got": float(v), "want": float(g["loss_" + k][0])} for k, v in ld.items()}
    params = list(model.parameters()) + ([t_emb_mod.weight] if args.beta else [])
    names = [n for n, _ in model.named_parameters()] + (["t_table"] if args.beta else [])
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    torch.cuda.synchronize()
    probe_gen = torch.Generator().manual_seed(99)
    gerr = {}
    for n, gr in zip(names, grads):
        key = "gradnorm_" + n
        if key not in g.files:
            gerr[n] = {"unexpected_grad": gr is not None and float(gr.abs().max()) > 0}
            continue
        want_norm, want_dot = g[key]
        shape = dict(zip(names, params))[n].shape
        probe = torch.randn(shape, generator=probe_gen).to(device)      # same stream as make_golden.py
        if gr is None:
            gerr[n] = {"missing": True}
            continue
        e = {"norm": float(gr.norm()), "want_norm": float(want_norm),
             "rel_dot_err": abs(float((gr * probe).sum()) - float(want_dot)) / (float(want_norm) + 1e-30),
             "nan": int(torch.isnan(gr).sum())}
        if "grad_" + n in g.files:
            w = torch.from_numpy(g["grad_" + n]).to(device)
            e["rel_l2"] = float((gr - w).norm() / (w.norm() + 1e-30))
        gerr[n] = e
    rep["grad"] = gerr
    return rep
