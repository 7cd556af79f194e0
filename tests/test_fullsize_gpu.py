"""Full-size parity (BASELINE.json configs 2 and 3 at their stated batch sizes) against the oracle evaluated in
fp32 on the same GPU, and the equivalence of the two product paths (the kernel-by-kernel `fused_step` that
bench.py's `value` times, and `render_rays` + loss classes + autograd that its `e2e` times).

The oracle's autograd graph over 8192 x 64 (or 16384 x 128) points does not fit a GPU in one piece, so its step
is evaluated the way any chunked reverse pass is: (1) forward over ray chunks without a graph, (2) the losses
and their gradients w.r.t. the per-ray outputs on the full batch (tiny), (3) forward + backward per ray chunk
with those upstream gradients, accumulating into the parameters.  Losses are sums over rays with global
denominators, so this is the full-batch gradient, not an approximation.

Tolerances (north_star): rgb max-abs 1e-3, depth 1e-2 m, parameter gradients full-matrix rel-L2 <= 1e-2."""
import types

import pytest
import torch

import spnerf_b200  # noqa: F401
from oracle import spnerf_oracle as O
from parity_common import TOL
from spnerf_b200 import config, synthetic, train_step
from spnerf_b200.models import load_model
from spnerf_b200.modules import metrics
from spnerf_b200.modules.rendering import render_rays

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(args):
    torch.manual_seed(0)
    model = load_model(args)
    with torch.no_grad():
        model.sigma_from_xyz[0].bias.fill_(3.0)
        model.sigma_from_xyz[0].weight.mul_(4.0)
    return model.to(DEV)


def _product_step(model, args, batch):
    res = render_rays({"coarse": model}, args, batch["rays"], None, semantics=batch["sems"], mode="train",
                      valid_depth=batch["valid_depth"], target_depths=batch["depths"], target_std=batch["depth_std"])
    loss = metrics.SNerfLoss(0.0)(res, batch["rgbs"])[0] \
        + metrics.DepthLoss(1.0, usealldepth=False)(res, batch["depths"][:, 0], batch["depths"][:, 1],
                                                    target_valid_depth=batch["valid_depth"],
                                                    target_std=batch["depth_std"])[0] \
        + metrics.SemanticLoss(1.0)(res, batch["sems"])[0]
    grads = torch.autograd.grad(loss, list(model.parameters()))
    return res, loss, grads


def _oracle_step_chunked(P, cfg, batch, z, z_unsort, chunk):
    """The oracle's forward, losses and parameter gradients on the given sample depths (see module docstring)."""
    rays = batch["rays"]
    b = rays.shape[0]
    o, d, sun = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]

    def run(sl):
        xyz = o[sl].unsqueeze(1) + d[sl].unsqueeze(1) * z[sl].unsqueeze(2)            # modules/rendering.py:147
        return O.inference(P, cfg, xyz, z[sl], sun[sl], batch["sems"][sl], None, None,
                           z_unsort=None if z_unsort is None else z_unsort[sl])
    keys = ("rgb", "depth", "weights", "sem_logits")
    with torch.no_grad():
        parts = [run(slice(i, i + chunk)) for i in range(0, b, chunk)]
    full = {k: torch.cat([p[k] for p in parts], 0) for k in keys}
    leaves = {k: full[k].clone().requires_grad_(True) for k in keys}
    res = {f"{k}_coarse": v for k, v in leaves.items()}
    res["z_vals_coarse"] = z
    loss = O.colour_loss(res, batch["rgbs"])[0] + O.depth_loss(
        res, batch["depths"][:, 0], batch["depths"][:, 1], batch["valid_depth"], batch["depth_std"], 1.0, False)[0] \
        + O.semantic_loss(res, batch["sems"], 1.0)[0]
    ups = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)))
    for p in P.values():
        p.grad = None
    for i in range(0, b, chunk):
        sl = slice(i, i + chunk)
        r = run(sl)
        outs = [r[k] for k in keys if ups[k] is not None]
        torch.autograd.backward(outs, [ups[k][sl] for k in keys if ups[k] is not None])
    return full, loss.detach(), {k: p.grad for k, p in P.items()}


def _compare(model, res, loss, grads, want, want_loss, want_grads):
    assert float((res["rgb_coarse"].detach() - want["rgb"]).abs().max()) <= TOL["rgb"]
    assert float((res["depth_coarse"].detach() - want["depth"]).abs().max()) <= TOL["depth"]
    assert float((res["weights_coarse"].detach() - want["weights"]).abs().max()) <= TOL["weights"]
    assert float((res["sem_logits_coarse"].detach() - want["sem_logits"]).abs().max()) <= TOL["sem_logits"]
    assert abs(float(loss) - float(want_loss)) <= 2e-3 * abs(float(want_loss))
    top = max(float(g.norm()) for g in want_grads.values())
    worst = {}
    for (name, _), g in zip(model.named_parameters(), grads):
        w = want_grads[name]
        assert bool(torch.isfinite(g).all()), name
        if float(w.norm()) < 1e-3 * top:
            continue                      # vanishing gradient: relative error is noise
        worst[name] = float((g - w).norm() / w.norm())
        assert worst[name] <= 1e-2, (name, worst[name])
    return worst


def test_c2_full_batch_against_the_oracle_on_the_same_gpu():
    """BASELINE config 2: 8192 rays x 64 samples, --depth --sem C=3: outputs, loss and EVERY parameter gradient
    (full matrices) of the product step against the oracle in fp32 on this GPU."""
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
    model = _model(args)
    batch = {k: v.to(DEV) for k, v in synthetic.make_batch(8192, seed=31).items()}
    torch.manual_seed(5)
    res, loss, grads = _product_step(model, args, batch)
    cfg = O.make_cfg(sem=True, num_sem_classes=3, fc_units=512)
    P = {k: v.detach().clone().requires_grad_(True) for k, v in model.named_parameters()}
    want, want_loss, want_grads = _oracle_step_chunked(P, cfg, batch, res["z_vals_coarse"], None, 2048)
    worst = _compare(model, res, loss, grads, want, want_loss, want_grads)
    print("c2 full size: worst gradient rel-L2", max(worst.items(), key=lambda kv: kv[1]))


def test_c3_full_batch_against_the_oracle_on_the_same_gpu():
    """BASELINE config 3: 16384 rays, --guidedsample --mapping (+ depth + sem): the second (128-sample) pass of the
    product against the oracle on the product's own merged depths (the sampler is pinned bit-exactly elsewhere)."""
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512, mapping=True, guidedsample=True, chunk=16384)
    model = _model(args)
    batch = {k: v.to(DEV) for k, v in synthetic.make_batch(16384, seed=32).items()}
    torch.manual_seed(6)
    res, loss, grads = _product_step(model, args, batch)
    z, z_unsort = res["z_vals_coarse"], res["z_vals_unsort_coarse"]
    assert z.shape == (16384, 128) and bool((z[:, 1:] >= z[:, :-1]).all())
    cfg = O.make_cfg(sem=True, num_sem_classes=3, fc_units=512, mapping=True, guidedsample=True)
    P = {k: v.detach().clone().requires_grad_(True) for k, v in model.named_parameters()}
    want, want_loss, want_grads = _oracle_step_chunked(P, cfg, batch, z, z_unsort, 1024)
    worst = _compare(model, res, loss, grads, want, want_loss, want_grads)
    print("c3 full size: worst gradient rel-L2", max(worst.items(), key=lambda kv: kv[1]))


class _FixedDraws:
    """args._rng object that replays the given uniform draws (one per call)."""

    def __init__(self, draws):
        self.draws = list(draws)

    def uniform(self, shape):
        t = self.draws.pop(0)
        assert tuple(t.shape) == tuple(shape)
        return t

    def normal(self, shape):          # drawn to keep the stream position; unused at noise_std = 0
        return torch.zeros(shape, device=self.draws[0].device if self.draws else DEV)


@pytest.mark.parametrize("rays", [8192, 1000])
def test_fused_step_equals_the_autograd_path(rays):
    """bench.py's `value` times train_step.fused_step; its `e2e` (and every other parity test) goes through
    render_rays + the loss classes + autograd.  Same draws -> same scalars and the same flat gradient."""
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
    model = _model(args)
    batch = {k: v.to(DEV) for k, v in synthetic.make_batch(rays, seed=33).items()}
    u = torch.rand(rays, args.n_samples, device=DEV, generator=torch.Generator(DEV).manual_seed(9))
    args._rng = _FixedDraws([u.clone()])
    res, loss, grads = _product_step(model, args, batch)
    want_flat = torch.cat([g.reshape(-1) for g in grads])
    want_terms = [float(metrics.SNerfLoss(0.0)(res, batch["rgbs"])[0]),
                  float(metrics.DepthLoss(1.0, usealldepth=False)(res, batch["depths"][:, 0], batch["depths"][:, 1],
                                                                 target_valid_depth=batch["valid_depth"],
                                                                 target_std=batch["depth_std"])[0]),
                  float(metrics.SemanticLoss(1.0)(res, batch["sems"])[0])]
    args._rng = _FixedDraws([u.clone()])
    flat, views, scalars, launches = train_step.fused_step(model, args, batch, repack=True)
    torch.cuda.synchronize()
    assert launches > 0
    got_terms = [float(x) for x in scalars[:3]]
    for a, w in zip(got_terms, want_terms):
        assert abs(a - w) <= 1e-6 * max(1.0, abs(w)), (got_terms, want_terms)
    assert flat.numel() == want_flat.numel()
    rel = float((flat - want_flat).norm() / want_flat.norm())
    assert rel <= 1e-5, rel
    for v, g in zip(views, grads):
        assert v.shape == g.shape


def test_no_grad_passes_save_no_activations():
    """render_rays under torch.no_grad() (validation, the first pass of guided sampling) must not write the
    ~13 KB per point of activation saves a training pass needs (needs_input_grad ignores grad mode)."""
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
    model = _model(args)
    batch = {k: v.to(DEV) for k, v in synthetic.make_batch(512, seed=34).items()}
    eng = model.engine
    with torch.no_grad():
        render_rays({"coarse": model}, args, batch["rays"], None, semantics=batch["sems"], mode="test")
    assert eng.last_saved_bytes == 0
    render_rays({"coarse": model}, args, batch["rays"], None, semantics=batch["sems"], mode="test")
    assert eng.last_saved_bytes == eng.save_bytes(512 * args.n_samples) > 0
    for p in model.parameters():
        p.requires_grad_(False)
    render_rays({"coarse": model}, args, batch["rays"], None, semantics=batch["sems"], mode="test")
    assert eng.last_saved_bytes == 0                  # frozen parameters: nothing to differentiate


def test_graphed_step_replays_the_fused_step():
    """train_step.GraphedStep (the step captured in one CUDA graph) at the reference's default batch of 1024 rays:
    same seed -> the replay reproduces the eager fused step (scalars bit-equal, gradient to atomics' round-off);
    consecutive replays draw different sample depths; a new batch is picked up through the static buffers."""
    from spnerf_b200 import engine as E
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
    model = _model(args)
    batch = {k: v.to(DEV) for k, v in synthetic.make_batch(1024, seed=35).items()}
    E.manual_seed(5)
    flat, _, scalars, launches = train_step.fused_step(model, args, batch, repack=True)
    want_flat, want_scalars = flat.clone(), scalars.clone()
    assert launches <= 11, launches
    gs = train_step.GraphedStep(model, args, batch)
    E.manual_seed(5)
    flat, _, scalars = gs()
    torch.cuda.synchronize()
    assert torch.equal(scalars[:3], want_scalars[:3])
    assert float((flat - want_flat).norm() / want_flat.norm()) <= 1e-5
    first = scalars.clone()
    flat, _, scalars = gs()                      # next replay: the device-side stream has advanced
    torch.cuda.synchronize()
    assert not torch.equal(scalars[:3], first[:3]) and bool(torch.isfinite(scalars[:3]).all())
    other = {k: v.to(DEV) for k, v in synthetic.make_batch(1024, seed=36).items()}
    E.manual_seed(5)
    flat2, _, sc2, _ = train_step.fused_step(model, args, other, repack=True)
    want2, wsc2 = flat2.clone(), sc2.clone()
    E.manual_seed(5)
    flat, _, scalars = gs(other)
    torch.cuda.synchronize()
    assert torch.equal(scalars[:3], wsc2[:3])
    assert float((flat - want2).norm() / want2.norm()) <= 1e-5


def test_two_gpu_allreduced_gradient_equals_the_mean_of_the_shards():
    """One process per GPU over NCCL (tests/manual/gpu2_allreduce_check.py under torchrun); skipped on a 1-GPU box."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`); the gloo world-size-2 test covers the host logic")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(root, "tests", "manual", "gpu2_allreduce_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(os.path.join(root, "gpurun_out", "allreduce_check.json")))
    assert res["ok"] and res["rel_l2"] <= 1e-6, res
