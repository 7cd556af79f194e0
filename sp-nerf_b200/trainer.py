"""Lightning-free training driver: mirror of NeRF_pl (main.py:19-176, 301-302) for the rebuilt path
(SURVEY 8f row 2).

What it keeps from the reference: the models dict (`coarse`, optional `t` embedding), the loss objects and their
gating (`ds_drop` / `ss_drop` as fractions of `max_train_steps`, `loss_without_beta` during the first two epochs),
`args.noise_std *= 0.9` every step, Adam(lr=args.lr, weight_decay=0) with StepLR(step_size=1, gamma=0.9) per
epoch (utils.py:317-318), and the epoch arithmetic of utils.get_epoch_number_from_train_step.
What changes: all parameters live in ONE flat fp32 buffer (the modules' parameters are views of it), the
gradient is one flat buffer, data parallelism is one all-reduce of that buffer (parallel.py), and the optimiser
is one fused kernel (spnerf_adam_step).  No dataset / RPC / logging code (out of scope).
"""
import types

import numpy as np
import torch

from . import engine as E
from . import parallel
from .models import load_model
from .modules import metrics
from .modules.rendering import render_rays


def get_parameters(models):
    """utils.get_parameters (utils.py:294-305)."""
    if isinstance(models, (list, tuple)):
        return [p for m in models for p in get_parameters(m)]
    if isinstance(models, dict):
        return [p for m in models.values() for p in get_parameters(m)]
    return list(models.parameters())


def flatten_parameters_(params):
    """Move the parameters into one contiguous fp32 buffer; every parameter becomes a view of it.
    Each view starts on a 16-byte boundary (the packers read parameters with vector loads); the padding
    words in between have zero gradient and stay zero.  Returns (flat buffer, offsets in floats)."""
    offsets, off = [], 0
    for p in params:
        offsets.append(off)
        off += (p.numel() + 3) & ~3
    flat = torch.zeros(off, dtype=torch.float32, device=params[0].device)
    for p, o in zip(params, offsets):
        view = flat[o:o + p.numel()].view(p.shape)
        view.copy_(p.data)
        p.data = view
    return flat, offsets


class Trainer:
    def __init__(self, args, device, n_train_rays=None, models=None):
        self.args = args
        self.device = torch.device(device)
        if models is None:
            models = {"coarse": load_model(args).to(self.device)}
            if args.beta:
                models["t"] = torch.nn.Embedding(getattr(args, "t_embbeding_vocab", 30),
                                                 args.t_embbeding_tau).to(self.device)
        self.models = models
        self.params = get_parameters(models)
        self.flat, self.offsets = flatten_parameters_(self.params)
        parallel.broadcast_(self.flat)               # replicas start from rank 0's weights, as under DDP
        if hasattr(models["coarse"], "engine") and self.device.type == "cuda":
            models["coarse"].engine.mark_dirty()
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grad = torch.zeros_like(self.flat)
        self.lr = float(args.lr)
        self.opt_steps = 0
        self.train_steps = 0
        self.n_train_rays = n_train_rays
        self.loss = metrics.load_loss(args)
        self.depth = bool(getattr(args, "depth", False))
        self.sem = bool(args.sem)
        max_steps = getattr(args, "max_train_steps", 500000)
        if self.depth:
            self.depth_loss = metrics.DepthLoss(lambda_ds=getattr(args, "ds_lambda", 0.0),
                                                GNLL=getattr(args, "GNLL", False),
                                                usealldepth=getattr(args, "usealldepth", False),
                                                margin=args.margin, stdscale=args.stdscale)
            self.ds_drop = np.round(getattr(args, "ds_drop", 0.25) * max_steps)            # main.py:31
        if self.sem:
            self.semantic_loss = metrics.SemanticLoss(lambda_ss=getattr(args, "ss_lambda", 4e-2))
            self.ss_drop = np.round(getattr(args, "ss_drop", 1) * max_steps)               # main.py:36
        self.use_ts = bool(args.beta)
        if args.beta:
            self.loss_without_beta = metrics.SNerfLoss(lambda_sc=args.sc_lambda)           # main.py:46

    # utils.get_epoch_number_from_train_step: one epoch = ceil(n_rays / batch_size) steps
    def get_current_epoch(self, tstep):
        if not self.n_train_rays:
            return 0
        per_epoch = int(np.ceil(self.n_train_rays / float(self.args.batch_size)))
        return int(tstep // per_epoch)

    def training_step(self, batch):
        """One optimisation step on a batch with the reference's keys (main.py:126-176).
        Returns (loss, loss_dict) like the pieces the reference logs."""
        a = self.args
        self.train_steps += 1
        epoch_before = self.get_current_epoch(self.train_steps - 1)
        rays, rgbs = batch["rays"], batch["rgbs"]
        ts = batch["ts"].squeeze() if self.use_ts else None
        sems = batch["sems"].squeeze() if self.sem else None
        results = render_rays(self.models, a, rays, ts, semantics=sems, mode='train',
                              valid_depth=batch["valid_depth"], target_depths=batch["depths"],
                              target_std=batch["depth_std"])
        if 'beta_coarse' in results and self.get_current_epoch(self.train_steps) < 2:
            loss, loss_dict = self.loss_without_beta(results, rgbs)
        else:
            loss, loss_dict = self.loss(results, rgbs)
        a.noise_std *= 0.9                                                                 # main.py:155
        if self.depth:
            d = batch["depths"]
            loss_depth, dd = self.depth_loss(results, d[:, 0], d[:, 1], target_valid_depth=batch["valid_depth"],
                                             target_std=batch["depth_std"])
            if self.train_steps < self.ds_drop:
                loss = loss + loss_depth
            loss_dict.update(dd)
        if self.sem:
            sem_loss, sd = self.semantic_loss(results, sems)
            if self.train_steps < self.ss_drop:
                loss = loss + sem_loss
            loss_dict.update(sd)
        grads = torch.autograd.grad(loss, self.params, allow_unused=True)
        for p, g, off in zip(self.params, grads, self.offsets):
            seg = self.grad[off:off + p.numel()]
            if g is None:
                seg.zero_()
            else:
                seg.copy_(g.reshape(-1))
        parallel.allreduce_mean_(self.grad)          # the one collective of a data-parallel step
        self.opt_steps += 1
        E.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.opt_steps, self.lr)
        self.models["coarse"].engine.mark_dirty()
        if self.get_current_epoch(self.train_steps) != epoch_before:
            self.lr *= 0.9                                                                 # StepLR(1, 0.9) per epoch
        return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}


    # -- checkpoints: the layout Lightning writes for NeRF_pl (main.py:19-59, 314-325): {"state_dict": {"nerf_coarse.<param>":
    # ..., "embedding_t.weight": ...}, "global_step": ...}; a reference checkpoint loads here and vice versa ------------
    _PREFIX = {"coarse": "nerf_coarse.", "t": "embedding_t."}

    def state_dict(self):
        out = {}
        for key, prefix in self._PREFIX.items():
            if key in self.models:
                for name, tensor in self.models[key].state_dict().items():
                    out[prefix + name] = tensor.detach().clone()
        return out

    def save_checkpoint(self, path):
        torch.save({"state_dict": self.state_dict(), "global_step": self.train_steps,
                    "epoch": self.get_current_epoch(self.train_steps), "lr": self.lr,
                    "optimizer_states": [{"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                                          "step": self.opt_steps}]}, path)

    def load_checkpoint(self, ckpt, strict=True):
        """`ckpt`: a path or an already loaded dict, from this trainer or from the reference's Lightning run
        (resume_from_checkpoint, main.py:325).  Parameters are copied INTO the flat buffer's views, so the optimiser
        and the packed operands keep working; optimiser moments are restored when the checkpoint is ours."""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location=self.device, weights_only=False)
        sd = ckpt.get("state_dict", ckpt)
        for key, prefix in self._PREFIX.items():
            if key not in self.models:
                continue
            part = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
            own = self.models[key].state_dict()
            missing, unexpected = sorted(set(own) - set(part)), sorted(set(part) - set(own))
            if strict and (missing or unexpected):
                raise RuntimeError(f"checkpoint does not match models['{key}']: missing {missing}, unexpected {unexpected}")
            with torch.no_grad():
                for name, dst in own.items():
                    if name in part:
                        if tuple(part[name].shape) != tuple(dst.shape):
                            raise RuntimeError(f"{prefix}{name}: shape {tuple(part[name].shape)} != {tuple(dst.shape)}")
                        dst.copy_(part[name])             # state_dict() tensors alias the parameters (views of self.flat)
        self.train_steps = int(ckpt.get("global_step", self.train_steps))
        opt = ckpt.get("optimizer_states")
        if opt and isinstance(opt[0], dict) and "exp_avg" in opt[0] and opt[0]["exp_avg"].numel() == self.flat.numel():
            self.exp_avg.copy_(opt[0]["exp_avg"])
            self.exp_avg_sq.copy_(opt[0]["exp_avg_sq"])
            self.opt_steps = int(opt[0].get("step", self.opt_steps))
        if "lr" in ckpt:
            self.lr = float(ckpt["lr"])
        if hasattr(self.models["coarse"], "engine") and self.device.type == "cuda":
            self.models["coarse"].engine.mark_dirty()

    # -- batches: the reference's DataLoader(shuffle=True, batch_size) over the pooled rays of all training images
    # (datasets/__init__.py, main.py:99-107), as one device-side permutation per epoch ------------------------------
    def epoch_batches(self, rays_pool, generator=None, drop_last=False):
        """`rays_pool`: dict of tensors with a common first dimension (the reference's batch keys), resident on the
        device.  Yields shuffled batches of args.batch_size rays (index_select on the device, no host round trip)."""
        n = next(iter(rays_pool.values())).shape[0]
        dev = next(iter(rays_pool.values())).device
        perm = torch.randperm(n, device=dev, generator=generator)
        bs = int(self.args.batch_size)
        for i in range(0, n, bs):
            idx = perm[i:i + bs]
            if drop_last and idx.numel() < bs:
                break
            yield {k: v.index_select(0, idx) for k, v in rays_pool.items()}
