"""Multi-GPU check (run under torchrun, 2+ ranks): the bucketed / overlapped gradient all-reduce of DoseTrainer gives
the same averaged gradients as one all-reduce after backward, and equals the mean of the per-rank gradients."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_manifest  # noqa: E402
from dose_prediction_b200 import networks, synth, training  # noqa: E402
from oracle import synth_ckpt  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = 64
man = [(k, ([1, (size // 16) ** 3, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
sd = synth_ckpt.make_state_dict(man, seed=0)
vol = synth.make_batch(1, size, seed=500 + rank)
x, gt = vol["dose_input"].cuda(), vol["gt"].cuda()


def grads(overlap):
    training.DDP_OVERLAP = overlap
    m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
    m.load_state_dict(sd)
    m.cuda().train()
    tr = training.DoseTrainer(m, 1, size)
    assert (tr.tail_off is not None) == overlap
    tr.forward_backward(x, gt)
    torch.cuda.synchronize()
    return tr.flat_g.clone(), tr


g_overlap, tr = grads(True)
g_plain, _ = grads(False)
# local (un-reduced) gradients: no process group for this trainer
training.DDP_OVERLAP = False
m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
m.load_state_dict(sd)
m.cuda().train()
tr_local = training.DoseTrainer(m, 1, size, process_group=None)
saved = training.allreduce_mean_
training.allreduce_mean_ = lambda flat, group=None: flat
tr_local.forward_backward(x, gt)
training.allreduce_mean_ = saved
torch.cuda.synchronize()
mean = tr_local.flat_g.clone()
dist.all_reduce(mean)
mean /= dist.get_world_size()
d1 = float((g_overlap - g_plain).abs().max())
d2 = float((g_overlap - mean).abs().max() / mean.abs().max())
if rank == 0:
    print("tail offset %d of %d; max |overlap - plain| = %.3e; max |overlap - mean(local)| / max = %.3e" % (
        tr.tail_off, tr.total, d1, d2), flush=True)
    assert d1 == 0.0 and d2 < 1e-6
    print("ddp check ok", flush=True)
dist.barrier()
dist.destroy_process_group()
