"""CPU oracle for the SP-NeRF ray-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch (fp32, no autocast, any device but meant for
CPU), the algorithm of the reference's hot path so that the CUDA product path can be checked
against it.  It is imported only by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs; nothing under sp-nerf_b200/ imports it.

Pinning: the reference (ShiningFeng/SP-NeRF) ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself: oracle/make_golden.py imports the
unmodified reference from /root/reference, injects the random draws, checks this restatement
against it bit for bit on CPU and writes tests/golden/*.npz (committed).  tests/test_oracle_golden.py
re-checks the oracle against those files everywhere.

Every function cites the reference lines it follows (paths relative to the reference root).
Parameters are passed as a dict with the reference's state_dict key names.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# configuration / random draws
# ------------------------------------------------------------------------------------------------
def make_cfg(**kw):
    """Fields read by the hot path (modules/opt.py:31-107 defaults where the reference has one)."""
    d = dict(n_samples=64, n_importance=0, model="sp-nerf", beta=False, guidedsample=False, sc_lambda=0.0,
             margin=1e-4, stdscale=1.0, chunk=5120, noise_std=0.0, num_sem_classes=3, s_embedding_factor=1,
             fc_layers=8, fc_units=512, mapping=False, t_embbeding_tau=4, sem=False, mapping_freqs=10,
             skips=(4,), siren=True)
    d.update(kw)
    return SimpleNamespace(**d)


class Draws:
    """Pre-generated random tensors consumed in the reference's call order:
    uniform(B,N) jitter, normal(B,N) noise, uniform(B,N) CDF draws, uniform(n_valid,N) GT draws,
    normal(B,2N), normal(B,2N)."""

    def __init__(self, uniforms=(), normals=()):
        self.uniforms = list(uniforms)
        self.normals = list(normals)

    def uniform(self, shape):
        t = self.uniforms.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t

    def normal(self, shape):
        t = self.normals.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t


# ------------------------------------------------------------------------------------------------
# point network  (models/spnerf.py:5-37 Mapping, :40-46 Siren, :273-369 SPNeRF.forward)
# ------------------------------------------------------------------------------------------------
def positional_encoding(x, n_freqs):
    """models/spnerf.py:18,32-37: for f in 2**linspace(0,n-1,n): [sin(f x), cos(f x)]; raw x is NOT kept."""
    bands = 2 ** torch.linspace(0, n_freqs - 1, n_freqs)
    parts = []
    for f in bands:
        parts.append(torch.sin(f * x))
        parts.append(torch.cos(f * x))
    return torch.cat(parts, -1)


def _lin(P, name, x):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


def point_network(P, cfg, xyz, sun_d, labels=None, t_emb=None):
    """models/spnerf.py:305-369.  Returns (n_points, 8 [+1 if beta] [+C if sem]) with columns
    [albedo(3), sigma, sun visibility, sky(3), (beta), (semantic logits)]."""
    enc = positional_encoding(xyz, cfg.mapping_freqs) if cfg.mapping else xyz          # :305
    if cfg.sem and labels is not None:                                                  # :308-319
        lab = labels.squeeze().long()
        lab = torch.where(lab == -100, torch.tensor(cfg.num_sem_classes, device=lab.device), lab)
        emb = F.embedding(lab, P["semantic_embedding.weight"], padding_idx=cfg.num_sem_classes)
        x_in = torch.cat((enc, emb), dim=1)
    else:
        x_in = enc
    siren = getattr(cfg, "siren", True)
    # `nl` of models/spnerf.py:178: Siren (sin(w0 x), w0 = 1) or ReLU for every hidden activation; with siren the
    # first trunk layer is Siren(w0=30) (:202)
    nl = torch.sin if siren else torch.relu
    h = x_in
    for i in range(cfg.fc_layers):                                                      # :325-329
        if i in cfg.skips:
            h = torch.cat([h, x_in], -1)
        h = _lin(P, f"fc_net.{2 * i}", h)
        h = torch.sin((30.0 if i == 0 else 1.0) * h) if siren else torch.relu(h)        # :40-46, :202
    sigma = F.softplus(_lin(P, "sigma_from_xyz.0", h))                                  # :333
    feats = _lin(P, "feats_from_xyz", h)                                                # :338
    rgb = torch.sigmoid(_lin(P, "rgb_from_xyzdir.2", nl(_lin(P, "rgb_from_xyzdir.0", feats))))
    rgb = rgb * (1 + 2 * 0.001) - 0.001                                                 # :346-347
    out = torch.cat([rgb, sigma], 1)
    s = torch.cat([feats, sun_d], -1)                                                   # :351-352
    s = nl(_lin(P, "sun_v_net.0", s))
    s = nl(_lin(P, "sun_v_net.2", s))
    s = nl(_lin(P, "sun_v_net.4", s))
    sun_v = torch.sigmoid(_lin(P, "sun_v_net.6", s))
    sky = torch.sigmoid(_lin(P, "sky_color.2", torch.relu(_lin(P, "sky_color.0", sun_d))))  # :355
    out = torch.cat([out, sun_v, sky], 1)
    if cfg.beta:                                                                        # :359-362
        b = torch.cat([feats, t_emb], -1)
        b = F.softplus(_lin(P, "beta_from_xyz.2", nl(_lin(P, "beta_from_xyz.0", b))))
        out = torch.cat([out, b], 1)
    if cfg.sem:                                                                         # :365-367
        lg = _lin(P, "logit_from_label.2", nl(_lin(P, "logit_from_label.0", h)))
        out = torch.cat([out, lg], 1)
    return out


def n_outputs(cfg):
    return 8 + (1 if cfg.beta else 0) + (cfg.num_sem_classes if cfg.sem else 0)        # :267-271


# ------------------------------------------------------------------------------------------------
# volume integration  (models/spnerf.py:63-159 inference)
# ------------------------------------------------------------------------------------------------
def composite(out, z, noise_std=0.0, noise=None, has_beta=False, has_sem=False):
    """models/spnerf.py:109-157 given the network output `out` (B,N,ncol) and depths z (B,N)."""
    albedo, sigma = out[..., :3], out[..., 3]
    sun_v, sky = out[..., 4:5], out[..., 5:8]
    d = z[:, 1:] - z[:, :-1]                                                            # :116-118
    d = torch.cat([d, 1e10 * torch.ones_like(d[:, :1])], -1)
    nz = (noise if noise is not None else torch.zeros_like(sigma)) * noise_std         # :121-122
    alpha = 1 - torch.exp(-d * torch.relu(sigma + nz))                                  # :123
    shifted = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1)         # :126
    trans = torch.cumprod(shifted, -1)[:, :-1]                                          # :127
    w = alpha * trans                                                                   # :128
    depth = torch.sum(w * z, -1)                                                        # :131
    irr = sun_v + (1 - sun_v) * sky                                                     # :132
    rgb = torch.clamp(torch.sum(w.unsqueeze(-1) * albedo * irr, -2), min=0., max=1.)    # :133-134
    res = {"rgb": rgb, "depth": depth, "weights": w, "transparency": trans, "albedo": albedo,
           "sun": sun_v, "sky": sky, "z_vals": z}
    idx = 8
    if has_beta:                                                                        # :148-152
        res["beta"] = out[..., idx:idx + 1]
        idx += 1
    if has_sem:                                                                         # :154-157 (plain mean)
        res["sem_logits"] = torch.mean(out[..., idx:], dim=1)
    return res


def inference(P, cfg, xyz, z, sun_d, labels=None, t_emb=None, draws=None, z_unsort=None):
    """models/spnerf.py:83-159: per-ray inputs repeated per sample, point-chunked network calls."""
    B, N = xyz.shape[0], xyz.shape[1]
    pts = xyz.reshape(-1, 3)
    sun_p = torch.repeat_interleave(sun_d, N, dim=0)                                    # :89
    lab_p = None if labels is None else torch.repeat_interleave(labels, N, dim=0)      # :91
    t_p = None if t_emb is None else torch.repeat_interleave(t_emb, N, dim=0)          # :90
    outs = []
    for i in range(0, pts.shape[0], cfg.chunk):                                         # :94-107
        outs.append(point_network(P, cfg, pts[i:i + cfg.chunk], sun_p[i:i + cfg.chunk],
                                  None if lab_p is None else lab_p[i:i + cfg.chunk],
                                  None if t_p is None else t_p[i:i + cfg.chunk]))
    out = torch.cat(outs, 0).view(B, N, n_outputs(cfg))
    noise = draws.normal((B, N)) if draws is not None else torch.randn(B, N, device=z.device)  # :122 (always drawn)
    res = composite(out, z, cfg.noise_std, noise, cfg.beta, cfg.sem)
    if z_unsort is not None:                                                            # :145-146
        res["z_vals_unsort"] = z_unsort
    return res


# ------------------------------------------------------------------------------------------------
# ray sampling  (modules/rendering.py:14-147)
# ------------------------------------------------------------------------------------------------
def stratified_z(rays, n, u):
    """modules/rendering.py:128-144 (use_disp False, perturb 1.0 hard-coded at :124-125)."""
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = torch.linspace(0, 1, n, device=rays.device)
    z = near * (1 - t) + far * t
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    hi = torch.cat([mid, z[:, -1:]], -1)
    lo = torch.cat([z[:, :1], mid], -1)
    return lo + (hi - lo) * (1.0 * u)


def inverse_cdf_draw(bins, weights, u, eps=1e-5):
    """modules/rendering.py:26-54 sample_pdf with det=False; u are the uniform draws (:35)."""
    n_bins = weights.shape[1]
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_bins)
    pair = torch.stack([below, above], -1).view(u.shape[0], -1)
    cdf_g = torch.gather(cdf, 1, pair).view(u.shape[0], -1, 2)
    bins_g = torch.gather(bins, 1, pair).view(u.shape[0], -1, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom[denom < eps] = 1
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0]), inds


def three_sigma_draw(lo, hi, n, near0, far0, u):
    """modules/rendering.py:58-73 sample_3sigma: Gaussian-weighted bins over [lo,hi] clamped to the
    near/far of the FIRST ray of the batch (rendering.py:95,113 pass near[0,0], far[0,0])."""
    t = torch.linspace(0., 1., steps=n, device=lo.device)
    step = (hi - lo) / (n - 1)
    edges = (lo.unsqueeze(-1) * (1. - t) + hi.unsqueeze(-1) * t).clamp(near0, far0)
    factor = (edges[..., 1:] - edges[..., :-1]) / step.unsqueeze(-1)
    x = torch.linspace(-3., 3., steps=(n - 1), device=lo.device)
    bw = factor * (1. / math.sqrt(2 * math.pi) * torch.exp(-0.5 * x.pow(2))).unsqueeze(0).expand(
        *edges.shape[:-1], n - 1)
    return inverse_cdf_draw(edges, bw, u)


def guided_z(res, z, n, near, far, mode, valid_depth, target_depths, target_std, draws, trace=None):
    """modules/rendering.py:76-116 GenerateGuidedSamples (+ compute_samples_around_depth).
    `trace` (tests only) receives the searchsorted indices of rendering.py:38 per ray."""
    depth, w = res["depth"], res["weights"]
    std = (((z - depth.unsqueeze(-1)).pow(2) * w).sum(-1)).sqrt()                       # :81
    z2, inds = three_sigma_draw(depth - 3. * std, depth + 3. * std, n, near[0, 0], far[0, 0],
                                draws.uniform((z.shape[0], n)))                         # :83-87
    if mode == "train":                                                                 # :98-114
        assert valid_depth is not None, 'valid_depth missing in training batch!'
        sel = valid_depth > 0
        td = torch.flatten(target_depths[:, 0][sel])
        ts = torch.flatten(target_std[sel])
        gt, inds_gt = three_sigma_draw(td - 3. * ts, td + 3. * ts, n, near[0, 0], far[0, 0],
                                       draws.uniform((int(sel.sum()), n)))
        z2[sel] = gt
        inds = inds.clone()
        inds[sel] = inds_gt
    if trace is not None:
        trace["inds"] = inds
        trace["std"] = std
    return z2


def render(P, cfg, rays, ts=None, labels=None, mode="test", valid_depth=None, target_depths=None,
           target_std=None, draws=None, t_table=None, trace=None):
    """modules/rendering.py:119-183 render_rays (coarse only; n_importance=0 in every config)."""
    if cfg.model != "sp-nerf":
        raise ValueError(f"model {cfg.model} is not valid")                             # :179
    n = cfg.n_samples
    o, d, near, far, sun_d = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    z = stratified_z(rays, n, draws.uniform((rays.shape[0], n)))                        # :131-144
    xyz = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)                              # :147
    t_emb = None
    if cfg.beta and ts is not None:                                                     # :155-156
        t_emb = F.embedding(ts, t_table)
    res = inference(P, cfg, xyz, z, sun_d, labels, t_emb, draws)                        # :157
    if cfg.guidedsample:                                                                # :159-170
        if trace is not None:
            trace.update(z1=z, weights1=res["weights"].detach(), depth1=res["depth"].detach())
        z2 = guided_z(res, z, n, near, far, mode, valid_depth, target_depths, target_std, draws, trace).detach()
        z2, _ = torch.sort(z2, -1)
        z_unsort = torch.cat([z, z2], -1)
        z, _ = torch.sort(torch.cat([z, z2], -1), -1)
        xyz = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
        res = inference(P, cfg, xyz, z, sun_d, labels, t_emb, draws, z_unsort=z_unsort)
    if cfg.sc_lambda > 0:                                                               # :171-177
        xyz_sc = o.unsqueeze(1) + sun_d.unsqueeze(1) * z.unsqueeze(2)
        tmp = inference(P, cfg, xyz_sc, z, sun_d, labels, t_emb, draws)
        res["weights_sc"] = tmp["weights"]
        res["transparency_sc"] = tmp["transparency"]
        res["sun_sc"] = tmp["sun"]
    return {f"{k}_coarse": v for k, v in res.items()}                                   # :181-183


# ------------------------------------------------------------------------------------------------
# losses  (modules/metrics.py:10-183)
# ------------------------------------------------------------------------------------------------
def solar_terms(res, lam):
    """modules/metrics.py:17-24."""
    sun = res["sun_sc_coarse"].squeeze()
    t2 = torch.sum(torch.square(res["transparency_sc_coarse"].detach() - sun), -1)
    t3 = 1 - torch.sum(res["weights_sc_coarse"].detach() * sun, -1)
    return {"coarse_sc_term2": lam / 3. * torch.mean(t2), "coarse_sc_term3": lam / 3. * torch.mean(t3)}


def colour_loss(res, target, lambda_sc=0.0, use_beta=False):
    """modules/metrics.py:27-45 SNerfLoss / :48-65 SatNerfLoss (+ :10-14 uncertainty_aware_loss)."""
    ld = {}
    if use_beta:
        beta = torch.sum(res["weights_coarse"].unsqueeze(-1) * res["beta_coarse"], -2) + 0.05
        ld["coarse_color"] = ((res["rgb_coarse"] - target) ** 2 / (2 * beta ** 2)).mean()
        ld["coarse_logbeta"] = (3 + torch.log(beta).mean()) / 2
    else:
        ld["coarse_color"] = F.mse_loss(res["rgb_coarse"], target, reduction="mean")
    if lambda_sc > 0:
        ld.update(solar_terms(res, lambda_sc))
    return sum(ld.values()), ld


def depth_loss(res, target_depth, target_weight, valid_depth, target_std, lambda_ds=1.0, usealldepth=False, gnll=False):
    """modules/metrics.py:68-159 DepthLoss: MSE variants and the GNLL subset variant (:76, :129-130: the
    predicted STD is passed where GaussianNLLLoss expects a variance; kept).  GNLL with usealldepth calls
    GaussianNLLLoss without a variance in the reference (:140) and raises there."""
    lam = lambda_ds / 3.                                                                # :71
    if usealldepth:                                                                     # :140, :154-156
        if gnll:
            raise TypeError("GaussianNLLLoss.forward() missing 1 required positional argument: 'var'")   # as :140 does
        per = (res["depth_coarse"] - target_depth) ** 2
        val = lam * torch.mean(target_weight * per)
        return val, {"coarse_ds": val}
    mask = valid_depth > 0                                                              # :89-92
    z, d, w = res["z_vals_coarse"][mask], res["depth_coarse"][mask], res["weights_coarse"][mask]
    if d.shape[0] == 0:                                                                 # :97-100
        per = torch.zeros((1,), device=target_weight.device, requires_grad=True)
    else:
        std = (((z - d.unsqueeze(-1)).pow(2) * w).sum(-1)).sqrt()                       # :102
        tw, td, tsd = target_weight[mask], target_depth[mask], target_std[mask]        # :105-107
        apply = torch.logical_or((d - td).abs() > tsd, std > tsd)                       # :78-80, :115
        d_a = d[apply]
        if d_a.shape[0] == 0:                                                           # :119-121
            per = torch.zeros((1,), device=target_weight.device, requires_grad=True)
        else:
            scale = float(d_a.shape[0]) / float(valid_depth.shape[0])                   # :125-127
            if gnll:                                                                    # :129-130
                per = scale * F.gaussian_nll_loss(d_a, td[apply], std[apply])
            else:
                per = scale * tw[apply] * (d_a - td[apply]) ** 2                        # :132
    val = lam * torch.mean(per)                                                         # :151-153
    return val, {"coarse_ds": val}


def semantic_loss(res, labels, lambda_ss=1.0):
    """modules/metrics.py:162-183: lambda * CE(ignore_index=-100, mean over kept rays)."""
    val = lambda_ss * F.cross_entropy(res["sem_logits_coarse"], labels, ignore_index=-100)
    return val, {"coarse_ss": val}


# ------------------------------------------------------------------------------------------------
# parameter construction with the reference's initialisation (models/spnerf.py:162-271)
# ------------------------------------------------------------------------------------------------
def input_width(cfg):
    base = 2 * cfg.mapping_freqs * 3 if cfg.mapping else 3                              # :185
    return base + (cfg.num_sem_classes * cfg.s_embedding_factor if cfg.sem else 0)      # :199


def parameter_shapes(cfg):
    """state_dict keys -> shapes, in registration order."""
    f, n_in, C = cfg.fc_units, input_width(cfg), cfg.num_sem_classes
    sh = {}
    if cfg.sem:
        sh["semantic_embedding.weight"] = (C + 1, C * cfg.s_embedding_factor)
    for i in range(cfg.fc_layers):
        k = n_in if i == 0 else (f + n_in if i in cfg.skips else f)
        sh[f"fc_net.{2 * i}.weight"], sh[f"fc_net.{2 * i}.bias"] = (f, k), (f,)
    sh["sigma_from_xyz.0.weight"], sh["sigma_from_xyz.0.bias"] = (1, f), (1,)
    sh["feats_from_xyz.weight"], sh["feats_from_xyz.bias"] = (f, f), (f,)
    if cfg.sem:
        sh["logit_from_label.0.weight"], sh["logit_from_label.0.bias"] = (f // 2, f), (f // 2,)
        sh["logit_from_label.2.weight"], sh["logit_from_label.2.bias"] = (C, f // 2), (C,)
    sh["rgb_from_xyzdir.0.weight"], sh["rgb_from_xyzdir.0.bias"] = (f // 2, f), (f // 2,)
    sh["rgb_from_xyzdir.2.weight"], sh["rgb_from_xyzdir.2.bias"] = (3, f // 2), (3,)
    sh["sun_v_net.0.weight"], sh["sun_v_net.0.bias"] = (f // 2, f + 3), (f // 2,)
    for j in (2, 4):
        sh[f"sun_v_net.{j}.weight"], sh[f"sun_v_net.{j}.bias"] = (f // 2, f // 2), (f // 2,)
    sh["sun_v_net.6.weight"], sh["sun_v_net.6.bias"] = (1, f // 2), (1,)
    sh["sky_color.0.weight"], sh["sky_color.0.bias"] = (f // 2, 3), (f // 2,)
    sh["sky_color.2.weight"], sh["sky_color.2.bias"] = (3, f // 2), (3,)
    if cfg.beta:
        t = cfg.t_embbeding_tau
        sh["beta_from_xyz.0.weight"], sh["beta_from_xyz.0.bias"] = (f // 2, f + t), (f // 2,)
        sh["beta_from_xyz.2.weight"], sh["beta_from_xyz.2.bias"] = (1, f // 2), (1,)
    return sh


def random_parameters(cfg, seed=0, sigma_bias=None):
    """Seeded stand-in weights with the reference's magnitudes (SIREN ranges for fc_net / sun_v_net,
    U(+-1/sqrt(fan_in)) elsewhere).  NOT the reference's RNG stream; tests that need the reference's
    exact init build the module instead.  sigma_bias shifts the density head so the transmittance
    scan is exercised ("trained-like")."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shape in parameter_shapes(cfg).items():
        fan_in = shape[-1] if len(shape) > 1 else None
        if name == "semantic_embedding.weight":
            t = torch.randn(shape, generator=g)
            t[cfg.num_sem_classes] = 0
        elif name.endswith(".bias"):
            w_shape = parameter_shapes(cfg)[name[:-5] + ".weight"]
            bound = 1 / math.sqrt(w_shape[-1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            if name.startswith(("fc_net.", "sun_v_net.")):
                first = name in ("fc_net.0.weight", "sun_v_net.0.weight")
                bound = 1 / fan_in if first else math.sqrt(6 / fan_in)
            else:
                bound = 1 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        P[name] = t
    if sigma_bias is not None:
        P["sigma_from_xyz.0.bias"] = torch.full((1,), float(sigma_bias))
    return P
