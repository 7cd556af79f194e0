"""Full-image inference (BASELINE config 4): the chunk loop of main.py:58-79 plus the per-pixel products
eval.py:60-101 forms for image export, kept on the device.

The reference renders a test image by calling render_rays on `args.chunk` rays at a time and moves EVERY
entry of the result dictionary to the host (main.py:75-76: weights, transparency, albedo, sun, sky, z_vals ...
about 5.6 KB per ray), then eval.py multiplies weights with sun / albedo / sky / beta per sample and takes the
argmax of the semantic logits, all on the CPU.  Here the compositing kernel emits those per-ray sums itself
(SpnerfCompositeFwd.ray_aux / sem_argmax), nothing per-sample is stored, and a rank ships 64
bytes per ray.  Rays are independent, so a multi-GPU render is a contiguous split of the image rows across
ranks with one gather of the per-ray outputs at the end; there is no collective on the data path.

Quirks kept: the coarse sampler is stochastic in test mode (Q1), guided sampling clamps to
the near/far of the first ray of each chunk (Q5).
"""
import torch

from . import engine as E
from . import parallel
from .modules import rendering as R


def _pass(model, args, rays, z, rays_t, labels, want_samples):
    """One forward pass without gradient bookkeeping; per-ray outputs (+ weights for the guided sampler)."""
    eng = model.engine
    eng.ensure_packed()
    n = z.shape[1]
    noise_std = float(args.noise_std)
    noise = torch.randn(z.shape, dtype=torch.float32, device=z.device) if noise_std != 0.0 else None
    sky, _ = eng.sky(rays)
    out, _ = eng.forward(rays, n, z=z, labels=labels, t_emb=rays_t, sky=sky, save=False)
    return E.composite_fwd(out, z, eng.n_out, eng.col_sem, eng.n_sem, noise=noise, noise_std=noise_std,
                           want_raw=False, want_samples=want_samples, want_aux=True,
                           col_beta=8 if model.beta else -1)


@torch.no_grad()
def render_chunk(models, args, rays, ts=None, semantics=None):
    """Per-ray outputs of one chunk of rays: the `test`-mode path of render_rays (modules/rendering.py:119-218)."""
    if args.model != "sp-nerf":
        raise ValueError(f'model {args.model} is not valid')
    if args.n_importance > 0:
        raise NotImplementedError("fine model (n_importance > 0) is outside the rebuilt path (SURVEY Q9)")
    model = models["coarse"]
    rays = rays.float().contiguous()
    E._require_cuda(rays, "rays")
    b, n, dev = rays.shape[0], args.n_samples, rays.device
    rng = getattr(args, "_rng", None)
    z = E.sample_coarse(rays, R._draw_uniform(rng, (b, n), dev), n)
    rays_t = None
    if args.beta and ts is not None:
        rays_t = models['t'](ts).detach().float().contiguous()
    labels = None
    if model.sem and semantics is not None:
        labels = semantics.detach().reshape(-1).long().contiguous()
    if args.guidedsample:
        w, _, _, _, depth, _, _, _ = _pass(model, args, rays, z, rays_t, labels, True)
        _, z = R.guided_depths({"weights": w, "depth": depth}, z, rays, 'test', None, None, None, rng)
    _, _, rgb, _, depth, sem, aux, cls = _pass(model, args, rays, z, rays_t, labels, False)
    res = {"rgb": rgb, "depth": depth, "albedo": aux[:, 0:3], "sun": aux[:, 3:4], "sky": aux[:, 4:7]}
    if model.beta:
        res["beta"] = aux[:, 7:8]
    if model.sem:
        res["sem_logits"] = sem
        res["sem_class"] = cls
    return res


@torch.no_grad()
def render_image(models, args, rays, ts=None, semantics=None, chunk=None):
    """All rays of an image on this GPU, `chunk` rays at a time (default args.chunk, main.py:60).
    Returns {rgb (B,3), depth (B), albedo (B,3), sun (B,1), sky (B,3), [beta (B,1)], [sem_logits (B,C),
    sem_class (B) int32]} on the device: rgb/depth/sem_logits are render_rays' `*_coarse` entries, the rest
    are eval.py:75-101's weighted sums."""
    chunk = int(chunk or getattr(args, "chunk", 0) or rays.shape[0])
    parts = []
    for i in range(0, rays.shape[0], chunk):
        sl = slice(i, i + chunk)
        parts.append(render_chunk(models, args, rays[sl], None if ts is None else ts[sl],
                                  None if semantics is None else semantics[sl]))
    if not parts:
        raise ValueError("render_image: no rays")
    return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}


@torch.no_grad()
def render_image_sharded(models, args, rays, ts=None, semantics=None, chunk=None, dst=0):
    """Multi-GPU render: every rank renders a contiguous block of the image's rays (parallel.shard_bounds) and
    rank `dst` receives the concatenated per-ray outputs (None elsewhere).  `rays` is the full (B,11) tensor on
    every rank (or at least this rank's block addressed by the same indices)."""
    rank, world = parallel.rank_world()
    lo, hi = parallel.shard_bounds(rays.shape[0], rank, world)
    local = render_image(models, args, rays[lo:hi], None if ts is None else ts[lo:hi],
                         None if semantics is None else semantics[lo:hi], chunk=chunk)
    out = {}
    for k in sorted(local):
        g = parallel.gather_rays(local[k], rays.shape[0], dst=dst)
        if g is not None:
            out[k] = g
    return out if rank == dst or world == 1 else None
