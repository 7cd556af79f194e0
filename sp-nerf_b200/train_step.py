"""One fused training step of the hot path, kernel by kernel, without the autograd machinery:
sampling -> point network -> volume integration -> losses -> adjoints -> parameter gradients.

This is the path bench.py times (inputs resident in HBM) and the core of a Lightning-free
trainer.  It is numerically the same computation as
``render_rays(...)`` + ``SNerfLoss`` + ``DepthLoss`` + ``SemanticLoss`` + ``loss.backward()``
(main.py:125-174 with --depth --sem), minus Python graph bookkeeping.
"""
import torch

from . import engine as E


class StepTimer:
    """Optional CUDA-event timing of each kernel group on the launching stream."""

    def __init__(self, enabled):
        self.enabled = enabled
        self.marks = []

    def mark(self, name):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def durations_ms(self):
        out = {}
        for (_, e0), (name, e1) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


def fused_step(model, args, batch, lambda_ds=1.0, lambda_ss=1.0, use_all_depth=False, repack=True, timer=None,
               allreduce=None):
    """batch: dict with the reference's keys (rays, rgbs, sems, valid_depth, depths, depth_std), on the device.
    Returns (flat gradient buffer, per-parameter views, loss scalars (8,) on the device, launches).
    `allreduce(flat)` is called on the flat gradient buffer when given (data parallel)."""
    if args.guidedsample or args.sc_lambda > 0 or args.beta:
        raise NotImplementedError("fused_step covers the --depth --sem configuration; use render_rays otherwise")
    eng = model.engine
    t = timer or StepTimer(False)
    rays = batch["rays"]
    b, n = rays.shape[0], args.n_samples
    labels = batch["sems"].reshape(-1) if model.sem else None
    launches = 0
    t.mark("start")
    eng.ensure_packed(force=repack)            # parameters change every optimiser step
    launches += 1 if repack else 0          # one pack launch: weight tiles (fwd / bwd), small fp32 block, aux tiles
    t.mark("pack")
    rng = getattr(args, "_rng", None)          # parity tests replay the reference's draws
    z = E.sample_coarse(rays, rng.uniform((b, n)), n) if rng is not None else E.sample_coarse_rng(rays, n)
    sky, sky_hidden = eng.sky(rays)
    launches += 2
    t.mark("sample")
    out, saves = eng.forward(rays, n, z=z, labels=labels, sky=sky, save=True)
    launches += 1
    t.mark("mlp_fwd")
    weights, trans, rgb, rgb_raw, depth, sem = E.composite_fwd(out, z, eng.n_out, eng.col_sem, eng.n_sem)
    launches += 1
    t.mark("composite_fwd")
    scalars, g_rgb, g_depth, g_sem, _ = E.losses(
        b, rgb=rgb, rgb_target=batch["rgbs"], depth=depth, z=z, weights=weights,
        target_depth=batch["depths"][:, 0], target_weight=batch["depths"][:, 1],
        target_std=batch["depth_std"], valid_depth=batch["valid_depth"], lambda_ds=lambda_ds,
        use_all_depth=use_all_depth, sem_logits=sem, labels=labels, lambda_ss=lambda_ss)
    launches += 2 if (sem is not None and sem.numel() > (1 << 18)) else 1      # large batches count the labelled rays in a launch of their own
    t.mark("losses")
    g_out, g_sky_ray, absmax = E.composite_bwd(out, z, weights, trans, rgb_raw, eng.n_out, eng.col_sem, eng.n_sem,
                                               g_rgb=g_rgb, g_depth=g_depth, g_sem=g_sem, absmax=eng.absmax)
    launches += 1
    t.mark("composite_bwd")
    flat, views, _ = eng.backward(g_out, out, rays, n, saves, absmax, labels=labels, g_sky_ray=g_sky_ray, sky=sky,
                                  sky_hidden=sky_hidden, timer=t, copy=False)
    launches += 4          # backward-data, sky backward, weight GEMMs, reduce + flush
    if allreduce is not None:
        allreduce(flat)
        t.mark("allreduce")
    return flat, views, scalars, launches


class GraphedStep:
    """`fused_step` captured once in a CUDA graph and replayed: one graph launch per training step instead of ~12
    kernel launches plus their Python / ctypes bookkeeping, which is what a step costs at the reference's default
    batch of 1024 rays (modules/opt.py:35) once the kernels themselves take ~1.5 ms.

    The batch lives in static device buffers (`load` copies a new batch in, from pinned host memory or the device);
    the sampler draws its uniforms inside the kernel from a device-resident counter, so every replay sees fresh
    random numbers; the weight repack is part of the graph, so the optimiser may update the parameters in place
    between replays.  Outputs (`flat`, `views`, `scalars`) are the same buffers on every replay.  A data-parallel
    all-reduce stays outside the graph: call it on `flat` after `replay()`."""

    def __init__(self, model, args, example_batch, lambda_ds=1.0, lambda_ss=1.0, use_all_depth=False, warmup=2):
        if getattr(args, "_rng", None) is not None:
            raise ValueError("GraphedStep draws on the device; injected draws (args._rng) cannot be captured")
        self.model, self.args = model, args
        dev = example_batch["rays"].device
        E._require_cuda(example_batch["rays"], "batch")
        self.static = {k: v.detach().clone() for k, v in example_batch.items()}
        kw = dict(lambda_ds=lambda_ds, lambda_ss=lambda_ss, use_all_depth=use_all_depth, repack=True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # one-time table uploads and attribute calls happen here
            for _ in range(max(1, warmup)):
                fused_step(model, args, self.static, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.flat, self.views, self.scalars, self.launches = fused_step(model, args, self.static, **kw)

    def load(self, batch):
        for k, buf in self.static.items():
            buf.copy_(batch[k], non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.flat, self.views, self.scalars

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        return self.replay()
