"""Round-2 side measurements (one B200): (a) what an inference-plan rebuild costs after the parameters changed (the
train -> validate loop of the reference), (b) the freeze=False training step at BASELINE configs[3]'s size, eager and as
CUDA graphs."""
import json
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

out = {}
dev = torch.device("cuda:0")
size = 128
tokens = (size // 16) ** 3
man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
sd = synth_ckpt.make_state_dict(man, seed=0)

# ---- (a) plan rebuild: eval forward (build), in-place parameter change, eval forward (rebuild), eval forward (cached)
model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
model.load_state_dict(sd, strict=True)
model.to(dev).eval()
x = synth.make_batch(8, size, seed=1)["dose_input"].to(dev)
t = []
for step in range(3):
    if step == 1:
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(1.0)                      # bumps every version counter: the cached plan is stale
    torch.cuda.synchronize(); t0 = time.time()
    y = model(x)
    torch.cuda.synchronize(); t.append(time.time() - t0)
out["dose_plan_batch8_128"] = {"first_build_s": t[0], "rebuild_after_parameter_change_s": t[1], "cached_forward_s": t[2],
                               "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}
print(json.dumps(out), flush=True)
del model, y
torch.cuda.empty_cache()

# ---- (b) training step, freeze=True vs freeze=False
vol = synth.make_batch(2, size, seed=1234)
xb, gt = vol["dose_input"].to(dev), vol["gt"].to(dev)
for freeze in (True, False):
    model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    tr = DoseTrainer(model, 2, size, freeze=freeze)
    for _ in range(2):
        loss = tr.step(xb, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss = tr.step(xb, gt)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 5
    l_eager = float(loss)
    tr.capture()
    for _ in range(2):
        loss = tr.step(xb, gt)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        loss = tr.step(xb, gt)
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 5
    tr.P.check_device_errors()
    out["train_freeze_%s" % freeze] = {"eager_ms": eager, "graph_ms": graph, "samples_per_s_graph": 2e3 / graph, "launches": len(tr.P.steps),
                                        "loss_after_eager": l_eager, "loss_after_graph": float(loss),
                                        "plan_GiB": tr.P.bytes_alloc / 2 ** 30, "health": tr.check_health()}
    print(json.dumps(out["train_freeze_%s" % freeze]), flush=True)
    del tr, model
    torch.cuda.empty_cache()
with open(os.path.join(ROOT, "gpurun_out", "r3_misc.json"), "w") as f:
    json.dump(out, f, indent=1)
