"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
which exists only in the build container) with injected random draws, and check the oracle
restatement (oracle/spnerf_oracle.py) against it bit for bit on CPU.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py          # rewrites tests/golden/, prints the pinning report

The reference needs `kornia.losses.ssim` only for an eval metric outside the hot path
(modules/metrics.py:7,210-215), so a stub module is registered before the import.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import spnerf_oracle as O  # noqa: E402


def import_reference():
    k = types.ModuleType("kornia")
    kl = types.ModuleType("kornia.losses")
    kl.ssim = lambda *a, **kw: None
    k.losses = kl
    sys.modules.setdefault("kornia", k)
    sys.modules.setdefault("kornia.losses", kl)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "modules"
                 or m.startswith("modules.")]:
        del sys.modules[name]
    sys.path.insert(0, REF)
    try:
        import models as ref_models
        from modules import rendering as ref_rendering, metrics as ref_metrics
    finally:
        sys.path.remove(REF)
    return ref_models, ref_rendering, ref_metrics


class InjectedRNG:
    """Replace torch.rand / rand_like / randn (looked up at call time by the reference,
    rendering.py:35,143 and spnerf.py:122) by pops from pre-generated lists."""

    def __init__(self, draws):
        self.d = draws

    def __enter__(self):
        self.saved = (torch.rand, torch.rand_like, torch.randn)
        torch.rand = lambda *s, **kw: self.d.uniform(tuple(s[0]) if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.rand_like = lambda t, **kw: self.d.uniform(tuple(t.shape))
        torch.randn = lambda *s, **kw: self.d.normal(tuple(s[0]) if len(s) == 1 and not isinstance(s[0], int) else s)
        return self

    def __exit__(self, *a):
        torch.rand, torch.rand_like, torch.randn = self.saved


def make_draws(seed, B, N, guided, sc, n_valid, train):
    g = torch.Generator().manual_seed(seed)
    uni = [torch.rand(B, N, generator=g)]
    nor = [torch.randn(B, N, generator=g)]
    if guided:
        uni.append(torch.rand(B, N, generator=g))
        if train:
            uni.append(torch.rand(n_valid, N, generator=g))
        nor.append(torch.randn(B, 2 * N, generator=g))
    if sc:
        nor.append(torch.randn(B, 2 * N if guided else N, generator=g))
    return uni, nor


def state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


CASES = {
    # BASELINE config 1 shape, shrunk in rays: fwd, test mode, --sem C=3, no mapping
    "c1_test_sem": dict(B=96, mode="test", seed=11, cfg=dict(sem=True, num_sem_classes=3, fc_units=512)),
    # BASELINE config 2 shape: train step with --depth --sem
    "c2_train_depth_sem": dict(B=96, mode="train", seed=12, noise_std=0.0, trained_like=True,
                               cfg=dict(sem=True, num_sem_classes=3, fc_units=512)),
    # BASELINE config 3 shape: --guidedsample --mapping (+ sem, + solar correction as in the README recipe)
    "c3_train_guided_mapping_sc": dict(B=64, mode="train", seed=13, trained_like=True,
                                       cfg=dict(sem=True, num_sem_classes=3, fc_units=512, mapping=True,
                                                guidedsample=True, sc_lambda=0.1, noise_std=0.25)),
    # guided sampling in test mode (no GT overwrite), no sem
    "guided_test_nosem": dict(B=48, mode="test", seed=14, trained_like=True,
                              cfg=dict(sem=False, fc_units=512, guidedsample=True)),
    # beta head + transient embedding, small trunk (stored with its weights)
    "beta_small": dict(B=32, mode="train", seed=15, store_weights=True,
                       cfg=dict(sem=True, num_sem_classes=5, fc_units=64, beta=True, mapping=True, sc_lambda=0.05)),
    # Sat-NeRF branch at the width the library builds: uncertainty head + transient embedding (models/spnerf.py:359-362),
    # SatNerfLoss with solar correction (metrics.py:10-24,48-65), --mapping, sem C=3
    "beta_512": dict(B=64, mode="train", seed=16, trained_like=True,
                     cfg=dict(sem=True, num_sem_classes=3, fc_units=512, beta=True, mapping=True, sc_lambda=0.05,
                              t_embbeding_tau=4)),
    # ReLU variant of the network (models/spnerf.py:178 `nl`): not reachable through load_model, built from the class
    "relu_512": dict(B=64, mode="train", seed=17, siren=False,
                     cfg=dict(sem=True, num_sem_classes=3, fc_units=512, mapping=True, siren=False)),
}
# parameters whose FULL gradient is stored whatever their size (two trunk-sized matrices; the rest of the
# large ones are pinned by norm + a random probe)
FULL_GRADS = ("fc_net.8.weight", "feats_from_xyz.weight")
FULL_GRAD_CASES = ("c2_train_depth_sem", "beta_512")


def run_case(name, spec, ref_models, ref_rendering, ref_metrics):
    from spnerf_b200 import synthetic
    B, mode = spec["B"], spec["mode"]
    cfg = O.make_cfg(**spec["cfg"])
    args = types.SimpleNamespace(**vars(cfg))
    args.chunk = 5120
    N = cfg.n_samples
    batch = synthetic.make_batch(B, seed=spec["seed"])
    train = mode == "train"
    n_valid = int((batch["valid_depth"] > 0).sum())

    torch.manual_seed(0)
    if spec.get("siren", True):
        ref_model = ref_models.load_model(args)
    else:
        ref_model = ref_models.SPNeRF(num_sem_classes=args.num_sem_classes, s_embedding_factor=args.s_embedding_factor,
                                      layers=args.fc_layers, feat=args.fc_units, mapping=args.mapping,
                                      t_embedding_dims=args.t_embbeding_tau, beta=args.beta, sem=args.sem, siren=False)
    t_table = None
    models = {"coarse": ref_model}
    if cfg.beta:
        models["t"] = torch.nn.Embedding(30, cfg.t_embbeding_tau)
        t_table = models["t"].weight
    if spec.get("trained_like"):
        with torch.no_grad():
            ref_model.sigma_from_xyz[0].bias.fill_(3.0)
            ref_model.sigma_from_xyz[0].weight.mul_(4.0)
    P = {k: v for k, v in ref_model.named_parameters()}

    uni, nor = make_draws(spec["seed"] + 1000, B, N, cfg.guidedsample, cfg.sc_lambda > 0, n_valid, train)
    kw = dict(semantics=batch["sems"] if cfg.sem else None, mode=mode)
    if train:
        kw.update(valid_depth=batch["valid_depth"], target_depths=batch["depths"], target_std=batch["depth_std"])
    ts = batch["ts"] if cfg.beta else None

    # ---- reference ----
    with InjectedRNG(O.Draws([u.clone() for u in uni], [n.clone() for n in nor])):
        ref_out = ref_rendering.render_rays(models, args, batch["rays"], ts, **kw)
    loss_fn = ref_metrics.load_loss(args)
    ref_loss, ref_ld = loss_fn(ref_out, batch["rgbs"])
    if train:
        dl = ref_metrics.DepthLoss(lambda_ds=1.0, GNLL=False, usealldepth=False)
        l2, d2 = dl(ref_out, batch["depths"][:, 0], batch["depths"][:, 1], target_valid_depth=batch["valid_depth"],
                    target_std=batch["depth_std"])
        ref_loss = ref_loss + l2
        ref_ld.update(d2)
    if cfg.sem:
        sl = ref_metrics.SemanticLoss(lambda_ss=1.0)
        l3, d3 = sl(ref_out, batch["sems"])
        ref_loss = ref_loss + l3
        ref_ld.update(d3)
    params = list(ref_model.parameters()) + ([t_table] if t_table is not None else [])
    ref_grads = torch.autograd.grad(ref_loss, params, allow_unused=True)

    # ---- oracle ----
    trace = {}
    ora_out = O.render(P, cfg, batch["rays"], ts, batch["sems"] if cfg.sem else None, mode,
                       batch["valid_depth"] if train else None, batch["depths"] if train else None,
                       batch["depth_std"] if train else None, O.Draws([u.clone() for u in uni], [n.clone() for n in nor]),
                       t_table=t_table, trace=trace)
    ora_loss, ora_ld = O.colour_loss(ora_out, batch["rgbs"], cfg.sc_lambda, cfg.beta)
    if train:
        l2, d2 = O.depth_loss(ora_out, batch["depths"][:, 0], batch["depths"][:, 1], batch["valid_depth"],
                              batch["depth_std"], 1.0, False)
        ora_loss = ora_loss + l2
        ora_ld.update(d2)
    if cfg.sem:
        l3, d3 = O.semantic_loss(ora_out, batch["sems"], 1.0)
        ora_loss = ora_loss + l3
        ora_ld.update(d3)
    ora_grads = torch.autograd.grad(ora_loss, params, allow_unused=True)

    # ---- pin: bit-for-bit ----
    report = {}
    assert set(ref_out) == set(ora_out), (sorted(ref_out), sorted(ora_out))
    for k in ref_out:
        assert torch.equal(ref_out[k], ora_out[k]), f"{name}: output {k} differs"
    for k in ref_ld:
        assert torch.equal(ref_ld[k].reshape(-1), ora_ld[k].reshape(-1)), f"{name}: loss {k} differs"
    for (pn, _), gr, go in zip(list(ref_model.named_parameters()) + ([("t", 0)] if t_table is not None else []),
                               ref_grads, ora_grads):
        assert (gr is None) == (go is None), pn
        if gr is not None:
            assert torch.equal(gr, go), f"{name}: grad {pn} differs"
    report["outputs_bit_equal"] = sorted(ref_out)
    report["losses_bit_equal"] = sorted(ref_ld)

    # ---- store ----
    store = {f"in_{k}": v.numpy() for k, v in batch.items()}
    for i, u in enumerate(uni):
        store[f"uniform_{i}"] = u.numpy()
    for i, n_ in enumerate(nor):
        store[f"normal_{i}"] = n_.numpy()
    for k, v in ref_out.items():
        store[f"out_{k}"] = v.detach().numpy()
    for k, v in ref_ld.items():
        store[f"loss_{k}"] = v.detach().reshape(-1).numpy()
    # sampler intermediates (first-pass weights / depth and the searchsorted indices): inputs and
    # expected outputs of the guided sampler taken in isolation, where bit-exactness is required
    for k, v in trace.items():
        store[f"mid_{k}"] = v.detach().numpy().astype(np.int32 if k == "inds" else np.float32)
    names = [n_ for n_, _ in ref_model.named_parameters()] + (["t_table"] if t_table is not None else [])
    g = torch.Generator().manual_seed(99)
    for n_, gr in zip(names, ref_grads):
        if gr is None:
            continue
        probe = torch.randn(gr.shape, generator=g)
        store[f"gradnorm_{n_}"] = np.array([float(gr.norm()), float((gr * probe).sum())], dtype=np.float64)
    # full gradients only for the small tensors (heads' last layers, biases, embedding)
    for n_, gr in zip(names, ref_grads):
        if gr is not None and (gr.numel() <= 1024 or (n_ in FULL_GRADS and name in FULL_GRAD_CASES)):
            store[f"grad_{n_}"] = gr.numpy()
    meta = dict(name=name, B=B, mode=mode, cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(cfg).items()},
                trained_like=bool(spec.get("trained_like")), state_sha256=state_hash(ref_model.state_dict()),
                n_valid=n_valid, torch=torch.__version__)
    if spec.get("store_weights"):
        for k, v in ref_model.state_dict().items():
            store[f"w_{k}"] = v.numpy()
        store["w_t_table"] = t_table.detach().numpy()
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **store)
    report["file_kb"] = os.path.getsize(os.path.join(out_dir, name + ".npz")) // 1024
    return report


def main():
    torch.set_num_threads(8)
    ref_models, ref_rendering, ref_metrics = import_reference()
    rep = {}
    for name, spec in CASES.items():
        rep[name] = run_case(name, spec, ref_models, ref_rendering, ref_metrics)
        print(name, rep[name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "PINNING.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
