"""Timing aid: OAR-TRANSEG training step (reference config: 4 crops of 96^3 per step)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dose_prediction_b200 import networks, synth  # noqa: E402
from dose_prediction_b200.training import SegTrainer  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
model = networks.OARTranseg(1, 8, (size,) * 3, pos_embed="perceptron").cuda().train()
tr = SegTrainer(model, batch, size)
vol = synth.make_batch(batch, size, seed=7)
ct, lab = vol["ct"].cuda(), synth.oar_labels(vol["oars"]).cuda()
for _ in range(3):
    loss = tr.step(ct, lab)
torch.cuda.synchronize()
tr.P.check_device_errors()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    loss = tr.step(ct, lab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("seg training step %dx%d^3: %.1f ms -> %.1f crops/s, loss %.4f, %d launches, %.1f GB" % (
    batch, size, ms, batch / ms * 1e3, float(loss), len(tr.P.steps), tr.P.bytes_alloc / 1e9))
fam = tr.P.profile_families()
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print("  %-24s %7.2f ms %4d launches" % (k, d["ms"], d["launches"]))
