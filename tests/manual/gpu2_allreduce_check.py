"""Data-parallel gradient equivalence on real GPUs (all-reduced gradients == single-GPU gradients
on the union batch).  Launch with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/manual/gpu2_allreduce_check.py

Every rank renders its own contiguous shard of a 2 x 4096-ray batch (fused step, NCCL all-reduce of the flat gradient
buffer), then recomputes every shard's gradient locally without the collective; the all-reduced buffer must equal the
mean of the per-shard gradients to 1e-6 (relative L2; float atomics in a few small slots reorder sums at ~1e-8).
Also checks that the Trainer's construction-time broadcast makes differently seeded replicas identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import spnerf_b200  # noqa: F401
from spnerf_b200 import config, engine as E, parallel, synthetic, train_step
from spnerf_b200.models import load_model
from spnerf_b200.trainer import Trainer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
    torch.manual_seed(0)
    model = load_model(args)
    with torch.no_grad():
        model.sigma_from_xyz[0].bias.fill_(3.0)
        model.sigma_from_xyz[0].weight.mul_(4.0)
    model = model.to(dev)
    per = 4096
    full = {k: v.to(dev) for k, v in synthetic.make_batch(per * world, seed=77).items()}
    mine = parallel.shard_batch(full, rank, world)
    assert mine["rays"].shape[0] == per

    E.manual_seed(100 + rank)
    flat, _, scalars, _ = train_step.fused_step(model, args, mine, repack=True, allreduce=parallel.allreduce_mean_)
    reduced = flat.clone()
    acc = torch.zeros_like(reduced)
    for s in range(world):
        E.manual_seed(100 + s)
        g, _, _, _ = train_step.fused_step(model, args, parallel.shard_batch(full, s, world), repack=True)
        acc += g
    want = acc / world
    rel = float((reduced - want).norm() / want.norm())
    max_abs = float((reduced - want).abs().max())
    # every rank holds the same reduced buffer
    copies = [torch.empty_like(reduced) for _ in range(world)]
    dist.all_gather(copies, reduced)
    same = all(torch.equal(c, copies[0]) for c in copies)

    torch.manual_seed(1000 + rank)                      # different initial weights per rank ...
    tr = Trainer(config.make_args(sem=True, num_sem_classes=3, fc_units=512, lr=5e-4, depth=True, ds_lambda=1.0), dev)
    w = [torch.empty_like(tr.flat) for _ in range(world)]
    dist.all_gather(w, tr.flat)
    synced = all(torch.equal(x, w[0]) for x in w)       # ... are rank 0's after construction
    loss, _ = tr.training_step(mine)
    w2 = [torch.empty_like(tr.flat) for _ in range(world)]
    dist.all_gather(w2, tr.flat)
    still = all(torch.equal(x, w2[0]) for x in w2) and not torch.equal(w2[0], w[0])

    res = {"world": world, "rays_per_rank": per, "rel_l2": rel, "max_abs": max_abs, "identical_on_all_ranks": same,
           "trainer_replicas_synced_at_construction": synced, "trainer_replicas_identical_after_a_step": still,
           "backend": dist.get_backend(), "ok": rel <= 1e-6 and same and synced and still}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "allreduce_check.json"), "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()
    if not res["ok"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
