"""One rendering pass (network on every sample + volume integration) as autograd functions over
the C ABI.  Filled in progressively; see models/spnerf.py and modules/rendering.py for the
reference-shaped entry points."""
import torch


def _rays_from_points(xyz, sun_d):
    """Per-point call surface -> the kernel's ray form: one sample per 'ray', origin = point."""
    n = xyz.shape[0]
    rays = torch.zeros(n, 11, dtype=torch.float32, device=xyz.device)
    rays[:, 0:3] = xyz
    rays[:, 8:11] = sun_d
    return rays


def point_rows(model, xyz, sun_d, t_emb=None, labels=None):
    eng = model.engine
    rays = _rays_from_points(xyz.float(), sun_d.float())
    lab = None
    if model.sem and labels is not None:
        lab = labels.reshape(-1).long().contiguous()
    out, _ = eng.forward(rays, 1, xyz=xyz.float().contiguous(), labels=lab,
                         t_emb=None if t_emb is None else t_emb.float().contiguous())
    return out


def integrate(model, args, z_vals, xyz=None, sun_d=None, rays_t=None, semantics=None, z_vals_unsort=None):
    raise NotImplementedError
