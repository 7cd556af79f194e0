"""One rendering pass (point network on every sample + volume integration) as a single autograd
node over the C ABI, plus the per-point call surface of SPNeRF.forward.

Forward: sky net -> fused MLP (saves activations when a backward will follow) -> compositing.
Backward: compositing adjoint -> fused MLP backward (data, then weights) -> sky net backward.
"""
import torch

from . import engine as E


class _Pass(torch.autograd.Function):
    """inputs: (module, rays, z, xyz, dir_override, labels, t_emb, noise, noise_std, grad_mode, *parameters)
    outputs: out (P,n_out), weights (B,N), transparency (B,N), rgb (B,3), depth (B), sem_logits (B,C)|empty
    `grad_mode` is torch.is_grad_enabled() at the call site: inside forward() grad mode is always off and
    needs_input_grad only mirrors the inputs' requires_grad flags, so neither tells a no_grad caller apart."""

    @staticmethod
    def forward(ctx, module, rays, z, xyz, dir_override, labels, t_emb, noise, noise_std, grad_mode, *params):
        eng = module.engine
        eng.ensure_packed()
        n = z.shape[1]
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)     # no activation saves for no_grad / frozen callers
        sky, sky_hidden = eng.sky(rays)
        t_emb_c = None if t_emb is None else t_emb.detach().float().contiguous()
        out, saves = eng.forward(rays, n, z=None if xyz is not None else z, xyz=xyz, dir_override=dir_override,
                                 labels=labels, t_emb=t_emb_c, sky=sky, save=need_grad)
        weights, trans, rgb, rgb_raw, depth, sem = E.composite_fwd(
            out, z, eng.n_out, eng.col_sem, eng.n_sem, noise=noise, noise_std=noise_std, want_raw=need_grad)
        ctx.module, ctx.n, ctx.noise_std = module, n, noise_std
        eng.last_saved_bytes = 0 if saves is None else saves.numel()      # observable by tests (no_grad => 0)
        ctx.has_t = t_emb is not None
        ctx.save_for_backward(rays, z, labels, t_emb_c, noise, out, saves, weights, trans, rgb_raw, sky, sky_hidden)
        if sem is None:
            sem = out.new_empty(0)
        return out, weights, trans, rgb, depth, sem

    @staticmethod
    def backward(ctx, g_out_ext, g_w, g_t, g_rgb, g_depth, g_sem):
        module, n = ctx.module, ctx.n
        eng = module.engine
        rays, z, labels, t_emb, noise, out, saves, weights, trans, rgb_raw, sky, sky_hidden = ctx.saved_tensors
        if saves is None:
            raise RuntimeError("backward through a rendering pass that was run without gradient tracking")

        def c(t):
            return None if t is None else t.float().contiguous()
        g_out, g_sky_ray, absmax = E.composite_bwd(
            out, z, weights, trans, rgb_raw, eng.n_out, eng.col_sem, eng.n_sem, g_rgb=c(g_rgb), g_depth=c(g_depth),
            g_sem=c(g_sem) if eng.n_sem > 0 and g_sem is not None and g_sem.numel() else None, g_w=c(g_w), g_t=c(g_t),
            g_out_ext=c(g_out_ext), noise=noise, noise_std=ctx.noise_std, absmax=eng.absmax)
        flat, views, g_temb = eng.backward(g_out, out, rays, n, saves, absmax, labels=labels, t_emb=t_emb,
                                           g_sky_ray=g_sky_ray, sky=sky, sky_hidden=sky_hidden)
        eng.last_grad_flat = flat          # the views below are slices of this one buffer (one all-reduce reaches them all)
        return (None, None, None, None, None, None, g_temb if ctx.has_t else None, None, None, None) + tuple(views)


def run_pass(module, rays, z, xyz=None, dir_override=None, labels=None, t_emb=None, noise=None, noise_std=0.0):
    params = tuple(module.parameters())
    return _Pass.apply(module, rays, z, xyz, dir_override, labels, t_emb, noise, float(noise_std),
                       torch.is_grad_enabled(), *params)


def _f32c(t):
    return t.detach().float().contiguous()


def integrate(module, args, z_vals, rays=None, xyz=None, sun_d=None, rays_t=None, semantics=None, z_vals_unsort=None,
              dir_override=None, noise=None):
    """Dictionary of one pass with the reference's keys (models/spnerf.py:136-157).
    Either `rays` (B,11) (points = origin + direction * z, never materialised) or explicit `xyz` (B,N,3)
    with `sun_d` (B,3), as in the reference's inference()."""
    b, n = z_vals.shape
    z = _f32c(z_vals)
    E._require_cuda(z, "z_vals")
    if rays is None:
        rays = torch.zeros(b, 11, dtype=torch.float32, device=z.device)
        rays[:, 8:11] = sun_d
        xyz = _f32c(xyz).reshape(-1, 3)
    else:
        rays = _f32c(rays)
        xyz = None
    labels = None
    if module.sem and semantics is not None:
        labels = semantics.detach().reshape(-1).long().contiguous()
    noise_std = float(args.noise_std)
    if noise is None and noise_std != 0.0:
        noise = torch.randn(b, n, dtype=torch.float32, device=z.device)        # models/spnerf.py:122
    out, weights, trans, rgb, depth, sem = run_pass(
        module, rays, z, xyz=xyz, dir_override=None if dir_override is None else _f32c(dir_override), labels=labels,
        t_emb=rays_t, noise=noise if noise_std != 0.0 else None, noise_std=noise_std)
    o3 = out.view(b, n, module.number_of_outputs)
    res = {"rgb": rgb, "depth": depth, "weights": weights, "transparency": trans, "albedo": o3[..., :3],
           "sun": o3[..., 4:5], "sky": o3[..., 5:8], "z_vals": z_vals}
    if z_vals_unsort is not None:
        res["z_vals_unsort"] = z_vals_unsort
    if module.beta:
        res["beta"] = o3[..., 8:9]
    if module.sem:
        res["sem_logits"] = sem
    return res


class _Rows(torch.autograd.Function):
    """SPNeRF.forward on explicit points: one 'ray' per point, no compositing."""

    @staticmethod
    def forward(ctx, module, rays, xyz, labels, t_emb, grad_mode, *params):
        eng = module.engine
        eng.ensure_packed()
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        sky, sky_hidden = eng.sky(rays)
        t_c = None if t_emb is None else t_emb.detach().float().contiguous()
        out, saves = eng.forward(rays, 1, xyz=xyz, labels=labels, t_emb=t_c, sky=sky, save=need_grad)
        ctx.module, ctx.has_t = module, t_emb is not None
        eng.last_saved_bytes = 0 if saves is None else saves.numel()
        ctx.save_for_backward(rays, labels, t_c, out, saves, sky, sky_hidden)
        return out

    @staticmethod
    def backward(ctx, g_out):
        module = ctx.module
        eng = module.engine
        rays, labels, t_c, out, saves, sky, sky_hidden = ctx.saved_tensors
        if saves is None:
            raise RuntimeError("backward through SPNeRF.forward that was run without gradient tracking")
        g_out = g_out.float().contiguous()
        absmax = g_out.abs().max().reshape(1)
        g_sky = g_out[:, 5:8].contiguous()
        _, views, g_t = eng.backward(g_out, out, rays, 1, saves, absmax, labels=labels, t_emb=t_c, g_sky_ray=g_sky,
                                     sky=sky, sky_hidden=sky_hidden)
        return (None, None, None, None, g_t if ctx.has_t else None, None) + tuple(views)


def point_rows(module, xyz, sun_d, t_emb=None, labels=None):
    xyz = _f32c(xyz)
    E._require_cuda(xyz, "input_xyz")
    n = xyz.shape[0]
    rays = torch.zeros(n, 11, dtype=torch.float32, device=xyz.device)
    rays[:, 8:11] = sun_d
    lab = None
    if module.sem and labels is not None:
        lab = labels.detach().reshape(-1).long().contiguous()
    return _Rows.apply(module, rays, xyz, lab, t_emb, torch.is_grad_enabled(), *tuple(module.parameters()))
