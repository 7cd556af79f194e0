"""Ray-sharded data parallelism: one process per GPU, the model replicated, every rank
renders its own contiguous block of rays, and the only collective is one all-reduce of the flat
fp32 parameter-gradient buffer per training step (NCCL over NVLink / NVSwitch on GPUs, gloo in the
CPU tests).  Full-image inference needs no collective until the final gather of per-ray outputs.

Denominators: every loss is a mean over the local shard (like DDP); after the all-reduce the
gradient is divided by the world size, i.e. the step optimises the mean of the per-shard losses.
The guided sampler's first-ray clamp is therefore per shard.
"""
import torch
import torch.distributed as dist


def rank_world():
    """(rank, world size) of this process; (0, 1) without a process group."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced block of `n_items` for `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_mean_(flat):
    """In-place mean over ranks of the flat gradient buffer (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda and dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)      # the division happens inside the collective: no extra launch
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)      # gloo has no AVG
            flat.div_(dist.get_world_size())
    return flat


def allreduce_grads_(module):
    """Mean over ranks of a module's .grad tensors after loss.backward().  A single-pass step leaves them as slices
    of the engine's flat gradient buffer (render_pass stores it as engine.last_grad_flat): one in-place all-reduce
    of that buffer reaches them all.  Otherwise (several passes summed by autograd) they are flattened first."""
    params = [p for p in module.parameters() if p.grad is not None]
    if not params or rank_world()[1] == 1:
        return
    flat = getattr(getattr(module, "_engine", None), "last_grad_flat", None)
    if flat is not None:
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        if all(lo <= p.grad.data_ptr() < hi for p in params) and sum(p.grad.numel() for p in params) == flat.numel():
            allreduce_mean_(flat)
            return
    grads = [p.grad for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    allreduce_mean_(flat)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)


def broadcast_(flat, src=0):
    """Every rank takes rank `src`'s copy of the flat parameter buffer (what DDP / Lightning do at construction:
    without it replicas built from different RNG states would stay different for the whole run)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def gather_rays(local, n_total, dst=0):
    """Concatenate per-ray outputs of a sharded inference on rank `dst` (others get None)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)
