"""Debug aid: one training step at small size on the GPU vs the CPU oracle; prints per-parameter gradient errors."""
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
from oracle import synth_ckpt, torch_ref  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tokens = (size // 16) ** 3
man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
sd = synth_ckpt.make_state_dict(man, seed=0)
vol = synth.make_batch(batch, size, seed=1234)
t0 = time.time()
probe = None
if os.environ.get("PROBE", "1") == "1":
    gen = torch.Generator().manual_seed(5)
    probe = [torch.randn((batch, 1) + (size >> i,) * 3, generator=gen) / (size >> i) ** 1.5 for i in range(4)]
if os.environ.get("EMU", "1") == "1":      # oracle forward with the CUDA path's operand rounding (same ReLU masks)
    from oracle import precision_probe
    torch_ref.EMU = precision_probe.recipe_dose
loss_ref, grads_ref, new_ref, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], probe=probe)
torch_ref.EMU = None
print("oracle train step %.1fs  loss %.6f" % (time.time() - t0, float(loss_ref)), flush=True)
model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
model.load_state_dict(sd, strict=True)
model.cuda().train()
tr = DoseTrainer(model, batch, size, lr=1e-4, weight_decay=1e-4, probe=probe)
print("plan: %d launches recorded, %.2f GB" % (len(tr.P.steps), tr.P.bytes_alloc / 1e9), flush=True)
loss = tr.forward_backward(vol["dose_input"].cuda(), vol["gt"].cuda())
torch.cuda.synchronize()
tr.P.check_device_errors()
print("loss gpu %.6f ref %.6f" % (float(loss), float(loss_ref)))
outs = tr.outputs()
for i, (a, b) in enumerate(zip(outs[1], outs_ref[1])):
    print("dose out %d rel_l2 %.3e" % (i, torch_ref.rel_l2(a.cpu(), b)))
g = tr.grads()
gmax = max(float(v.double().norm()) for v in grads_ref.values())
worst = []
for n in grads_ref:
    a, b = g[n].cpu().double(), grads_ref[n].double()
    err = float((a - b).norm())
    rel = err / max(float(b.norm()), 1e-30)
    worst.append((rel, n, float(b.norm()), float(a.norm())))
    print("%-70s rel %.3e  |ref| %.3e |got| %.3e" % (n, rel, float(b.norm()), float(a.norm())))
worst.sort(reverse=True)
print("WORST:")
for w in worst[:15]:
    print("  %.3e %s ref %.3e got %.3e" % w)
sd_new = {k: v.clone() for k, v in model.state_dict().items()}
bn = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.1."
print("running_mean rel %.3e  running_var rel %.3e" % (
    torch_ref.rel_l2(sd_new[bn + "running_mean"].cpu(), new_ref[bn + "running_mean"]),
    torch_ref.rel_l2(sd_new[bn + "running_var"].cpu(), new_ref[bn + "running_var"])))
tr2_loss = tr.step(vol["dose_input"].cuda(), vol["gt"].cuda())
torch.cuda.synchronize()
sd_new = model.state_dict()
dmax = 0.0
for n in grads_ref:
    d = float((sd_new[n].cpu() - new_ref[n]).abs().max())
    dmax = max(dmax, d)
print("max |param - ref| after one AdamW step: %.3e (lr 1e-4)" % dmax)
