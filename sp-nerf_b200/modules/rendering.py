"""Ray renderer: mirror of modules/rendering.py:119-218 (render_rays) over the sm_100a kernels.

Same call, same result-dictionary keys (suffix ``_coarse``), same quirks kept for parity
(SURVEY Appendix B): sampling is stochastic in test mode too (Q1), guided samples are clamped to the
near/far of the first ray of the batch (Q5), semantic logits are the plain mean over samples (Q3).
What changes is how it runs: the sample points are never materialised, the per-ray inputs are
never repeated per sample, the first pass of guided sampling runs without saving activations
(its graph is discarded by the reference as well, rendering.py:164,169), and there is no
device-to-host synchronisation (the reference has four per guided call, rendering.py:101-114).

Random draws: by default drawn on the device with torch.  For parity tests ``args._rng`` may hold
an object with ``uniform(shape)`` / ``normal(shape)`` that returns pre-generated tensors in the
reference's call order; the public signature is unchanged.
"""
import torch

from .. import engine as E
from .. import render_pass


def _draw_uniform(rng, shape, device):
    if rng is not None:
        return rng.uniform(shape).to(device=device, dtype=torch.float32).contiguous()
    return torch.rand(shape, dtype=torch.float32, device=device)


def _draw_normal(rng, shape, device, needed):
    if rng is not None:
        return rng.normal(shape).to(device=device, dtype=torch.float32).contiguous()   # keeps the stream position
    return torch.randn(shape, dtype=torch.float32, device=device) if needed else None


def guided_depths(res, z, rays, mode, valid_depth, target_depths, target_std, rng):
    """GenerateGuidedSamples + sort/concat/sort (modules/rendering.py:92-116, 165-167) on the device."""
    b, n = z.shape
    dev = z.device
    u_pred = _draw_uniform(rng, (b, n), dev)
    u_gt = None
    if mode == 'train':
        assert valid_depth is not None, 'valid_depth missing in training batch!'          # rendering.py:99
        valid_depth = valid_depth.reshape(-1).long().contiguous()
        if rng is not None:
            # the reference draws (n_valid, n) in compacted row order (rendering.py:113 via sample_pdf:35)
            sel = valid_depth > 0
            u_gt = torch.zeros(b, n, dtype=torch.float32, device=dev)
            u_gt[sel] = _draw_uniform(rng, (int(sel.sum()), n), dev)
        else:
            u_gt = torch.rand(b, n, dtype=torch.float32, device=dev)
        return E.sample_guided(rays, z, res["weights"].detach(), res["depth"].detach(), u_pred,
                               valid_depth=valid_depth, target_depths=target_depths.float(),
                               target_std=target_std.float().contiguous(), u_gt=u_gt)
    return E.sample_guided(rays, z, res["weights"].detach(), res["depth"].detach(), u_pred)


def render_rays(models, args, rays, ts, semantics=None, mode='test', valid_depth=None, target_depths=None,
                target_std=None):
    n = args.n_samples
    if args.model != "sp-nerf":
        raise ValueError(f'model {args.model} is not valid')                               # rendering.py:179
    if args.n_importance > 0:
        raise NotImplementedError("fine model (n_importance > 0) is outside the rebuilt path: every "
                                  "configuration of the reference uses n_importance = 0 (SURVEY Q9)")
    model = models["coarse"]
    rays = rays.float().contiguous()
    E._require_cuda(rays, "rays")
    rng = getattr(args, "_rng", None)
    b, dev = rays.shape[0], rays.device
    noisy = float(args.noise_std) != 0.0

    if rng is not None:                                                                    # rendering.py:131-144
        z = E.sample_coarse(rays, _draw_uniform(rng, (b, n), dev), n)
    else:
        z = E.sample_coarse_rng(rays, n)       # uniforms drawn inside the kernel (Philox), no torch.rand launch
    rays_t = None
    if args.beta:
        rays_t = models['t'](ts) if ts is not None else None                               # rendering.py:155-156
    common = dict(rays=rays, rays_t=rays_t, semantics=semantics)

    if args.guidedsample:                                                                  # rendering.py:159-170
        with torch.no_grad():   # the first pass only feeds the sampler
            first = render_pass.integrate(model, args, z, noise=_draw_normal(rng, (b, n), dev, noisy), **common)
        z_unsort, z = guided_depths(first, z, rays, mode, valid_depth, target_depths, target_std, rng)
        result = render_pass.integrate(model, args, z, z_vals_unsort=z_unsort,
                                       noise=_draw_normal(rng, (b, 2 * n), dev, noisy), **common)
    else:
        result = render_pass.integrate(model, args, z, noise=_draw_normal(rng, (b, n), dev, noisy), **common)

    if args.sc_lambda > 0:                                                                 # rendering.py:171-177
        tmp = render_pass.integrate(model, args, z, dir_override=rays[:, 8:11],
                                    noise=_draw_normal(rng, tuple(z.shape), dev, noisy), **common)
        result['weights_sc'] = tmp["weights"]
        result['transparency_sc'] = tmp["transparency"]
        result['sun_sc'] = tmp["sun"]
    return {f"{k}_coarse": v for k, v in result.items()}
