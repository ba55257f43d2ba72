"""Timing aid: DOSE-PYFER training step (config 4: batch 2 per GPU, 128^3) — step time and per-family breakdown."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_manifest  # noqa: E402
from dose_prediction_b200 import networks, synth  # noqa: E402
from dose_prediction_b200.training import DoseTrainer  # noqa: E402
from oracle import synth_ckpt  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tokens = (size // 16) ** 3
man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
sd = synth_ckpt.make_state_dict(man, seed=0)
model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
model.load_state_dict(sd, strict=True)
model.cuda().train()
t0 = time.time()
tr = DoseTrainer(model, batch, size, lr=1e-4, weight_decay=1e-4)
torch.cuda.synchronize()
print("plan built in %.1fs: %d launches, %.2f GB, %d refreshed weight tensors" % (
    time.time() - t0, len(tr.P.steps), tr.P.bytes_alloc / 1e9, len(tr.P.refresh)), flush=True)
vol = synth.make_batch(batch, size, seed=1234)
x, gt = vol["dose_input"].cuda(), vol["gt"].cuda()
for _ in range(2):
    loss = tr.step(x, gt)
torch.cuda.synchronize()
tr.P.check_device_errors()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps):
    loss = tr.step(x, gt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("step %.1f ms  -> %.2f samples/s  loss %.4f" % (ms, batch / ms * 1e3, float(loss)), flush=True)
e0.record()
tr.P.refresh_weights()
e1.record()
torch.cuda.synchronize()
print("refresh_weights %.1f ms" % e0.elapsed_time(e1))
fam = tr.P.profile_families()
tot = sum(d["ms"] for d in fam.values())
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    fl = tr.P.flops.get(k, 0.0)
    print("%-24s %8.2f ms %5d launches %s" % (k, d["ms"], d["launches"], ("%.0f TFLOP/s" % (fl / d["ms"] / 1e9)) if fl else ""))
print("sum of launches %.1f ms" % tot)
rows = tr.P.profile_launches()
rows.sort(key=lambda r: -r[2])
for n, l, ms_, f in rows[:25]:
    print("  %-20s %-46s %7.3f ms %s" % (n, l, ms_, ("%.0f TF/s" % (f / ms_ / 1e9)) if f else ""))
print("peak mem %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
