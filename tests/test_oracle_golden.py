"""The CPU oracle against the golden vectors produced from the unmodified reference
(oracle/make_golden.py).  Runs anywhere; pins the checker that the GPU parity tests rely on."""
import os
import numpy as np
import pytest
import torch

from oracle import spnerf_oracle as O
from parity_common import GOLDEN, build_model, load_case, make_args, state_hash

CASES = ["c1_test_sem", "c2_train_depth_sem", "c3_train_guided_mapping_sc", "guided_test_nosem", "beta_small", "beta_512", "relu_512"]


def _params(name, g, meta):
    if "w_t_table" in g.files:      # weights stored with the case
        P = {k[2:]: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in g.files
             if k.startswith("w_") and k != "w_t_table"}
        return P, torch.from_numpy(g["w_t_table"]).clone().requires_grad_(True)
    model, t_mod, _ = build_model(meta, "cpu")      # the transient table follows the model in the seeded stream
    assert state_hash(model.state_dict()) == meta["state_sha256"], "seeded init no longer matches the reference's"
    return dict(model.named_parameters()), (t_mod.weight if t_mod is not None else None)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g, meta = load_case(name)
    args = make_args(meta)
    cfg = O.make_cfg(**{k: v for k, v in vars(args).items()})
    P, t_table = _params(name, g, meta)
    ins = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    draws = O.Draws([torch.from_numpy(g[f"uniform_{i}"]) for i in range(10) if f"uniform_{i}" in g.files],
                    [torch.from_numpy(g[f"normal_{i}"]) for i in range(10) if f"normal_{i}" in g.files])
    train = meta["mode"] == "train"
    trace = {}
    out = O.render(P, cfg, ins["rays"], ins["ts"] if cfg.beta else None, ins["sems"] if cfg.sem else None,
                   meta["mode"], ins["valid_depth"] if train else None, ins["depths"] if train else None,
                   ins["depth_std"] if train else None, draws, t_table=t_table, trace=trace)
    assert sorted(out) == sorted(k[4:] for k in g.files if k.startswith("out_"))
    for k, v in out.items():
        want = torch.from_numpy(g["out_" + k])
        # same torch build on the same ISA reproduces the bits; leave room for a different host CPU
        assert torch.allclose(v.detach(), want, rtol=0, atol=5e-5), (k, float((v.detach() - want).abs().max()))
    if not cfg.guidedsample:
        assert torch.equal(out["z_vals_coarse"], torch.from_numpy(g["out_z_vals_coarse"]))
    loss, ld = O.colour_loss(out, ins["rgbs"], cfg.sc_lambda, cfg.beta)
    if train:
        l2, d2 = O.depth_loss(out, ins["depths"][:, 0], ins["depths"][:, 1], ins["valid_depth"], ins["depth_std"], 1.0,
                              False)
        loss, ld = loss + l2, {**ld, **d2}
    if cfg.sem:
        l3, d3 = O.semantic_loss(out, ins["sems"], 1.0)
        loss, ld = loss + l3, {**ld, **d3}
    for k, v in ld.items():
        assert abs(float(v) - float(g["loss_" + k][0])) <= 1e-5 * max(1.0, abs(float(g["loss_" + k][0]))), k
    names = list(P) + (["t_table"] if t_table is not None else [])
    tensors = list(P.values()) + ([t_table] if t_table is not None else [])
    grads = torch.autograd.grad(loss, tensors, allow_unused=True)
    for n, gr in zip(names, grads):
        if "gradnorm_" + n in g.files:
            assert gr is not None, n
            assert abs(float(gr.norm()) - g["gradnorm_" + n][0]) <= 1e-3 * g["gradnorm_" + n][0] + 1e-9, n


def test_sampler_recipe_matches_torch_reductions():
    """The guided sampler's bit-exactness rests on two facts about torch-CPU: sum over a
    contiguous fp32 row is an 8-lane vector accumulation, cumsum keeps a double running sum."""
    rng = np.random.default_rng(0)

    def row_sum(x):
        n, m = len(x), len(x) // 8
        acc4 = np.zeros((4, 8), np.float32)
        full = m // 4
        for i in range(full):
            for k in range(4):
                acc4[k] = acc4[k] + x[(4 * i + k) * 8:(4 * i + k) * 8 + 8]
        for v in range(4 * full, m):
            acc4[0] = acc4[0] + x[v * 8:v * 8 + 8]
        for k in range(1, 4):
            acc4[0] = acc4[0] + acc4[k]
        acc = np.float32(0)
        for j in range(8 * m, n):
            acc = np.float32(acc + x[j])
        for lane in range(8):
            acc = np.float32(acc + acc4[0][lane])
        return acc

    for n in (63, 64, 127, 128):
        for _ in range(200):
            x = rng.random(n).astype(np.float32)
            assert np.float32(row_sum(x)) == np.float32(torch.from_numpy(x.reshape(1, -1)).sum(-1).item())
    for _ in range(200):
        x = rng.random(63).astype(np.float32)
        x /= x.sum()
        run, e = 0.0, []
        for v in x:
            run += float(v)
            e.append(np.float32(run))
        assert np.array_equal(np.array(e, np.float32), torch.cumsum(torch.from_numpy(x.reshape(1, -1)), -1).numpy()[0])


def test_depth_loss_variants_against_the_reference():
    """tests/golden/depth_loss_variants.npz (oracle/make_golden_losses.py, the reference's DepthLoss): GNLL subset,
    MSE all-depth, MSE subset - loss values and gradients w.r.t. depth and weights."""
    import numpy as np
    g = np.load(os.path.join(GOLDEN, "depth_loss_variants.npz"))
    t = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    for name, kw in (("gnll_subset", dict(gnll=True, usealldepth=False)), ("mse_all", dict(usealldepth=True)),
                     ("mse_subset", dict(usealldepth=False))):
        d = t["depth"].clone().requires_grad_(True)
        w = t["weights"].clone().requires_grad_(True)
        res = {"z_vals_coarse": t["z"], "depth_coarse": d, "weights_coarse": w}
        val, _ = O.depth_loss(res, t["target_depth"], t["target_weight"], t["valid"], t["target_std"], lambda_ds=1.5, **kw)
        gd, gw = torch.autograd.grad(val, [d, w], allow_unused=True)
        assert abs(float(val) - float(g[name + "_loss"][0])) <= 1e-6 * max(1.0, abs(float(g[name + "_loss"][0]))), name
        assert torch.allclose(gd, torch.from_numpy(g[name + "_g_depth"]), rtol=1e-5, atol=1e-8), name
        if gw is not None:
            assert torch.allclose(gw, torch.from_numpy(g[name + "_g_weights"]), rtol=1e-5, atol=1e-8), name
        else:
            assert not g[name + "_g_weights"].any()
