"""Data-parallel plumbing on CPU: world_size 2 over gloo.  Sharding covers the batch exactly once,
the gradient all-reduce yields the mean of the per-rank buffers, the inference gather restores ray
order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import spnerf_b200
from spnerf_b200 import parallel, synthetic


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = synthetic.make_batch(101, seed=4)
        mine = parallel.shard_batch(batch, rank, world)
        flat = torch.full((1000,), float(rank + 1))
        parallel.allreduce_mean_(flat)
        gathered = parallel.gather_rays(mine["rays"][:, :3].clone(), 101, dst=0)
        ok = bool(torch.allclose(flat, torch.full((1000,), (1 + world) / 2)))
        ok = ok and parallel.rank_world() == (rank, world)
        lo, hi = parallel.shard_bounds(101, rank, world)
        cls = parallel.gather_rays(torch.arange(lo, hi, dtype=torch.int32), 101, dst=0)     # int32 per-ray output
        if rank == 0:
            ok = ok and torch.equal(gathered, batch["rays"][:, :3])
            ok = ok and cls.dtype == torch.int32 and torch.equal(cls, torch.arange(101, dtype=torch.int32))
        else:
            ok = ok and gathered is None
        # replicas built from different RNG states end up with rank 0's weights (DDP-style broadcast at construction)
        from spnerf_b200 import config
        from spnerf_b200.trainer import Trainer
        torch.manual_seed(100 + rank)
        tr = Trainer(config.make_args(sem=True, num_sem_classes=3, fc_units=512, lr=5e-4), "cpu")
        copies = [torch.empty_like(tr.flat) for _ in range(world)]
        dist.all_gather(copies, tr.flat)
        ok = ok and all(torch.equal(c, copies[0]) for c in copies) and float(tr.flat.abs().sum()) > 0
        ok = ok and all(p.data_ptr() >= tr.flat.data_ptr() for p in tr.params)      # still views of the flat buffer
        q.put((rank, ok, mine["rays"].shape[0]))
    finally:
        dist.destroy_process_group()


def test_allreduce_and_gather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert sum(r[2] for r in res) == 101
