"""Loss functions: mirror of modules/metrics.py:10-194 (same classes, constructor / forward
signatures and loss-dictionary keys) over the fused loss kernels of include/spnerf_b200.h.

Every term is evaluated by a CUDA kernel that yields the scalar(s) deterministically; gradients come
from the same kernels (colour MSE, depth, cross-entropy: value and gradient in one pass; solar
correction and uncertainty: a gradient pass scaled by the upstream gradients on the device).
The only PyTorch arithmetic left here is adding the scalars of a loss dictionary.
"""
import torch

from .. import engine as E


class _Reduce(torch.autograd.Function):
    """Wraps one fused loss evaluation: forward returns the scalar, backward scales the stored gradient."""

    @staticmethod
    def forward(ctx, which, x, kwargs):
        scalars, g_rgb, g_depth, g_sem, _ = E.losses(x.shape[0], **kwargs)
        grad = (g_rgb, g_depth, g_sem)[which]
        ctx.save_for_backward(grad)
        return scalars[which].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g * grad, None


def _c(t, dtype=torch.float32):
    return t.detach().to(dtype).contiguous()


def _f32(t):
    """fp32 view without forcing contiguity (the depth-loss kernel reads the two columns of `depths` in place)."""
    return t.detach() if t.dtype == torch.float32 else t.detach().float()


class _SolarTerms(torch.autograd.Function):
    """(sc_term2, sc_term3) of metrics.py:17-24 as one (2,) tensor; differentiable w.r.t. the sun visibility of the
    solar-correction pass only (the reference detaches the transparency and the weights)."""

    @staticmethod
    def forward(ctx, sun_sc, trans_sc, weights_sc, lambda_sc):
        trans_sc, weights_sc = _c(trans_sc), _c(weights_sc)
        ctx.lambda_sc = lambda_sc
        ctx.save_for_backward(sun_sc, trans_sc, weights_sc)
        return E.loss_solar(trans_sc, weights_sc, sun_sc, lambda_sc)

    @staticmethod
    def backward(ctx, g):
        sun_sc, trans_sc, weights_sc = ctx.saved_tensors
        g_sun = E.loss_solar(trans_sc, weights_sc, sun_sc, ctx.lambda_sc, upstream=_c(g), backward=True)
        return g_sun.view(sun_sc.shape), None, None, None


class _UncertaintyTerms(torch.autograd.Function):
    """(color, logbeta) of metrics.py:10-14 as one (2,) tensor; differentiable w.r.t. rgb, weights and beta."""

    @staticmethod
    def forward(ctx, rgb, weights, beta, gt_rgb, beta_min):
        rgb_c, w_c, gt = _c(rgb), _c(weights), _c(gt_rgb)
        vals, beta_ray = E.loss_uncertainty(rgb_c, gt, w_c, beta, beta_min)
        ctx.beta_min = beta_min
        ctx.save_for_backward(rgb_c, w_c, beta, gt, beta_ray)
        return vals

    @staticmethod
    def backward(ctx, g):
        rgb_c, w_c, beta, gt, beta_ray = ctx.saved_tensors
        g_rgb, g_w, g_beta = E.loss_uncertainty(rgb_c, gt, w_c, beta, ctx.beta_min, beta_ray=beta_ray, upstream=_c(g),
                                                backward=True)
        return g_rgb, g_w, g_beta.view(beta.shape), None, None


def solar_correction(loss_dict, inputs, typ, lambda_sc=0.05):
    """metrics.py:17-24: adds '<typ>_sc_term2' / '<typ>_sc_term3'."""
    terms = _SolarTerms.apply(inputs[f'sun_sc_{typ}'], inputs[f'transparency_sc_{typ}'], inputs[f'weights_sc_{typ}'],
                              float(lambda_sc))
    loss_dict[f'{typ}_sc_term2'], loss_dict[f'{typ}_sc_term3'] = terms[0], terms[1]
    return loss_dict


def uncertainty_aware_loss(loss_dict, inputs, gt_rgb, typ, beta_min=0.05):
    """metrics.py:10-14: adds '<typ>_color' / '<typ>_logbeta' (the transient uncertainty is always the coarse one)."""
    terms = _UncertaintyTerms.apply(inputs[f'rgb_{typ}'], inputs[f'weights_{typ}'], inputs['beta_coarse'], gt_rgb,
                                    float(beta_min))
    loss_dict[f'{typ}_color'], loss_dict[f'{typ}_logbeta'] = terms[0], terms[1]
    return loss_dict


class SNerfLoss(torch.nn.Module):
    """metrics.py:27-45."""

    def __init__(self, lambda_sc=0.05):
        super().__init__()
        self.lambda_sc = lambda_sc

    def forward(self, inputs, targets):
        rgb = inputs['rgb_coarse']
        loss_dict = {'coarse_color': _Reduce.apply(0, rgb, dict(rgb=_c(rgb), rgb_target=_c(targets)))}
        if self.lambda_sc > 0:
            loss_dict = solar_correction(loss_dict, inputs, 'coarse', self.lambda_sc)
        loss = sum(l for l in loss_dict.values())
        return loss, loss_dict


class SatNerfLoss(torch.nn.Module):
    """metrics.py:48-65."""

    def __init__(self, lambda_sc=0.0):
        super().__init__()
        self.lambda_sc = lambda_sc

    def forward(self, inputs, targets):
        loss_dict = uncertainty_aware_loss({}, inputs, targets, 'coarse')
        if self.lambda_sc > 0:
            loss_dict = solar_correction(loss_dict, inputs, 'coarse', self.lambda_sc)
        loss = sum(l for l in loss_dict.values())
        return loss, loss_dict


class _GnllDepth(torch.autograd.Function):
    """GNLL subset depth loss: value + gradients w.r.t. the depth and the weights (through the predicted STD)."""

    @staticmethod
    def forward(ctx, depth, weights, kwargs):
        scalars, _, g_depth, _, g_w = E.losses(depth.shape[0], **kwargs)
        ctx.save_for_backward(g_depth, g_w)
        return scalars[1].clone()

    @staticmethod
    def backward(ctx, g):
        g_depth, g_w = ctx.saved_tensors
        return g * g_depth, g * g_w, None


class DepthLoss(torch.nn.Module):
    """metrics.py:68-159.  One fused CUDA pass per call: the MSE variants (subset and all-depth) and the GNLL subset
    variant (metrics.py:76,129-130: torch's GaussianNLLLoss with the predicted STD passed as the variance, kept;
    nothing selected -> zero loss with zero gradients, :97-100,119-121)."""

    def __init__(self, lambda_ds=1.0, GNLL=False, usealldepth=True, margin=0, stdscale=1):
        super().__init__()
        self.lambda_ds = lambda_ds / 3.
        self._lambda_arg = lambda_ds
        self.GNLL, self.usealldepth, self.margin, self.stdscale = GNLL, usealldepth, margin, stdscale

    def forward(self, inputs, targets, weights=1., target_valid_depth=None, target_std=None):
        depth = inputs['depth_coarse']
        b = depth.shape[0]
        if self.GNLL:
            if self.usealldepth:       # the reference calls GaussianNLLLoss without a variance here (metrics.py:140)
                raise TypeError("GaussianNLLLoss.forward() missing 1 required positional argument: 'var' "
                                "(--GNLL needs the subset path, as in the reference)")
            kw = dict(depth=_c(depth), target_depth=_c(targets), target_weight=torch.ones(b, device=depth.device),
                      lambda_ds=self._lambda_arg, use_all_depth=False, z=_c(inputs['z_vals_coarse']),
                      weights=_c(inputs['weights_coarse']), target_std=_c(target_std), gnll=True,
                      valid_depth=None if target_valid_depth is None else _c(target_valid_depth, torch.int64))
            val = _GnllDepth.apply(depth, inputs['weights_coarse'], kw)
            return val, {'coarse_ds': val}
        if not torch.is_tensor(weights):
            weights = torch.full((b,), float(weights), device=depth.device)
        kw = dict(depth=_c(depth), target_depth=_f32(targets), target_weight=_f32(weights), lambda_ds=self._lambda_arg,
                  use_all_depth=self.usealldepth)
        if not self.usealldepth:
            kw.update(z=_c(inputs['z_vals_coarse']), weights=_c(inputs['weights_coarse']), target_std=_c(target_std),
                      valid_depth=None if target_valid_depth is None else _c(target_valid_depth, torch.int64))
        loss_dict = {'coarse_ds': _Reduce.apply(1, depth, kw)}
        loss = sum(l for l in loss_dict.values())
        return loss, loss_dict


class SemanticLoss(torch.nn.Module):
    """metrics.py:162-183."""

    def __init__(self, lambda_ss=1.0):
        super().__init__()
        self.lambda_ss = lambda_ss

    def forward(self, inputs, targets):
        logits = inputs['sem_logits_coarse']
        kw = dict(sem_logits=_c(logits), labels=_c(targets.reshape(-1), torch.int64), lambda_ss=self.lambda_ss)
        loss_dict = {'coarse_ss': _Reduce.apply(2, logits, kw)}
        loss = sum(loss_dict.values())
        return loss, loss_dict


def load_loss(args):
    """metrics.py:186-194."""
    if args.model != "sp-nerf":
        raise ValueError(f'model {args.model} is not valid')
    return SatNerfLoss(lambda_sc=args.sc_lambda) if args.beta else SNerfLoss(lambda_sc=args.sc_lambda)
