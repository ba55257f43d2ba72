"""Multi-GPU host logic on CPU: volume sharding is a partition, checked across 2 gloo ranks."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dose_prediction_b200.cascade import shard_volumes


def test_shards_partition_the_volumes():
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in shard_volumes(n, r, world))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_volumes(4, 2, 2)


def _worker(rank, world, port, n_volumes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_volumes(n_volumes, rank, world)
    owner = torch.zeros(n_volumes, dtype=torch.int64)
    owner[mine] = 1
    dist.all_reduce(owner)                       # test-only check: every volume has exactly one owner
    t = torch.tensor([float(len(mine))])
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        out.put((owner.tolist(), [int(g.item()) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    owner, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert owner == [1] * 9 and counts == [5, 4]


def _grad_worker(rank, world, port, out):
    from dose_prediction_b200.training import allreduce_mean_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(8, dtype=torch.float32) * (rank + 1)          # rank-dependent "gradients"
    allreduce_mean_(flat)
    if rank == 0:
        out.put(flat.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_is_the_mean():
    """training config (BASELINE configs[3]): data-parallel replicas exchange ONE flat gradient buffer."""
    from dose_prediction_b200.training import allreduce_mean_
    assert allreduce_mean_(torch.ones(3)).tolist() == [1.0, 1.0, 1.0]        # no process group: no-op
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 300
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [1.5 * i for i in range(8)]
