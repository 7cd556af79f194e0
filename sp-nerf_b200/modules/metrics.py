"""Loss functions: mirror of modules/metrics.py:10-194 (same classes, constructor / forward
signatures and loss-dictionary keys) over the fused loss kernel of include/spnerf_b200.h.

Colour MSE, depth supervision (subset and all-depth MSE variants) and semantic cross-entropy are
one CUDA pass each that yields the scalar and its gradient together; the solar-correction and
uncertainty terms (metrics.py:10-24) are a handful of elementwise ops on (rays, samples) tensors
and stay in PyTorch for now.
"""
import torch

from .. import engine as E


class _Reduce(torch.autograd.Function):
    """Wraps one fused loss evaluation: forward returns the scalar, backward scales the stored gradient."""

    @staticmethod
    def forward(ctx, which, x, kwargs):
        scalars, g_rgb, g_depth, g_sem = E.losses(x.shape[0], **kwargs)
        grad = (g_rgb, g_depth, g_sem)[which]
        ctx.save_for_backward(grad)
        return scalars[which].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g * grad, None


def _c(t, dtype=torch.float32):
    return t.detach().to(dtype).contiguous()


def solar_correction(loss_dict, inputs, typ, lambda_sc=0.05):
    """metrics.py:17-24."""
    sun_sc = inputs[f'sun_sc_{typ}'].squeeze()
    term2 = torch.sum(torch.square(inputs[f'transparency_sc_{typ}'].detach() - sun_sc), -1)
    term3 = 1 - torch.sum(inputs[f'weights_sc_{typ}'].detach() * sun_sc, -1)
    loss_dict[f'{typ}_sc_term2'] = lambda_sc / 3. * torch.mean(term2)
    loss_dict[f'{typ}_sc_term3'] = lambda_sc / 3. * torch.mean(term3)
    return loss_dict


def uncertainty_aware_loss(loss_dict, inputs, gt_rgb, typ, beta_min=0.05):
    """metrics.py:10-14."""
    beta = torch.sum(inputs[f'weights_{typ}'].unsqueeze(-1) * inputs['beta_coarse'], -2) + beta_min
    loss_dict[f'{typ}_color'] = ((inputs[f'rgb_{typ}'] - gt_rgb) ** 2 / (2 * beta ** 2)).mean()
    loss_dict[f'{typ}_logbeta'] = (3 + torch.log(beta).mean()) / 2
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


class DepthLoss(torch.nn.Module):
    """metrics.py:68-159.  The MSE variants (subset and all-depth) are one fused CUDA pass; the GNLL subset
    variant (metrics.py:76,129-130: torch's GaussianNLLLoss with the predicted STD passed as the variance, kept)
    is a handful of (rays, samples) torch ops on the device, like the solar-correction terms."""

    def __init__(self, lambda_ds=1.0, GNLL=False, usealldepth=True, margin=0, stdscale=1):
        super().__init__()
        self.lambda_ds = lambda_ds / 3.
        self._lambda_arg = lambda_ds
        self.GNLL, self.usealldepth, self.margin, self.stdscale = GNLL, usealldepth, margin, stdscale

    def _gnll_subset(self, inputs, targets, target_valid_depth, target_std):
        """ComputeSubsetDepthLoss with GNLL (metrics.py:82-130) without boolean-index gathers or host syncs:
        masked sums over all rays give the same mean over the selected ones."""
        z, depth, w = inputs['z_vals_coarse'].detach(), inputs['depth_coarse'], inputs['weights_coarse']
        E._require_cuda(depth, "depth_coarse")
        b = depth.shape[0]
        if target_valid_depth is None:
            target_valid_depth = torch.ones(b, device=depth.device)                       # metrics.py:86
        valid = target_valid_depth.reshape(-1) > 0
        pred_std = ((z - depth.unsqueeze(-1)).pow(2) * w).sum(-1).clamp_min(0).sqrt()      # :102
        apply = valid
        if not self.usealldepth:
            apply = valid & (((depth - targets).abs() > target_std) | (pred_std > target_std))     # :78-80, :115
        n_apply = apply.sum()
        var = torch.where(apply, pred_std, torch.ones_like(pred_std)).clamp_min(1e-6)      # GaussianNLLLoss eps
        per = 0.5 * (torch.log(var) + (depth - targets) ** 2 / var)
        mean = torch.where(apply, per, torch.zeros_like(per)).sum() / n_apply.clamp_min(1).to(per.dtype)
        scale = n_apply.to(per.dtype) / float(b)                                           # :125-127
        return self.lambda_ds * scale * mean       # zero (with zero gradient) when nothing is selected (:97-100,119-121)

    def forward(self, inputs, targets, weights=1., target_valid_depth=None, target_std=None):
        depth = inputs['depth_coarse']
        b = depth.shape[0]
        if self.GNLL:
            if self.usealldepth:       # the reference calls GaussianNLLLoss without a variance here (metrics.py:140)
                raise TypeError("GaussianNLLLoss.forward() missing 1 required positional argument: 'var' "
                                "(--GNLL needs the subset path, as in the reference)")
            val = self._gnll_subset(inputs, targets.float(), target_valid_depth, target_std.float())
            return val, {'coarse_ds': val}
        if not torch.is_tensor(weights):
            weights = torch.full((b,), float(weights), device=depth.device)
        kw = dict(depth=_c(depth), target_depth=_c(targets), target_weight=_c(weights), lambda_ds=self._lambda_arg,
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


def mse(image_pred, image_gt, valid_mask=None, reduction='mean'):
    """metrics.py:197-204."""
    value = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        value = value[valid_mask]
    return torch.mean(value) if reduction == 'mean' else value


def psnr(image_pred, image_gt, valid_mask=None, reduction='mean'):
    """metrics.py:206-207."""
    return -10 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))
