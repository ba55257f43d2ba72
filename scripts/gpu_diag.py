"""GPU diagnostic: eager-vs-eager determinism and eager-vs-CUDA-graph replay, buffer by buffer."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dose_prediction_b200 import networks, synth  # noqa: E402


def snapshot(P):
    return [t.clone() for t in P.keep if isinstance(t, torch.Tensor)]


def first_diff(a, b, label):
    n = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x.shape != y.shape or x.dtype != y.dtype:
            continue
        if not torch.equal(x, y):
            xf, yf = x.double(), y.double()
            rel = float((xf - yf).norm() / yf.norm().clamp_min(1e-30))
            if n < 6:
                print(f"  [{label}] buffer {i} shape {tuple(x.shape)} {x.dtype} differs rel={rel:.3e}")
            n += 1
    print(f"  [{label}] {n} differing buffers of {len(a)}")


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    seg = networks.OARTranseg(1, 8, (size,) * 3, pos_embed="perceptron").eval().to(dev)
    vol = synth.make_batch(2, size, seed=3)
    x = vol["ct"].to(dev)
    P = networks.plan_oar_transeg(seg, x.shape, dev)
    P.x_in.copy_(x)
    P.run(); torch.cuda.synchronize(); s1 = snapshot(P)
    P.run(); torch.cuda.synchronize(); s2 = snapshot(P)
    first_diff(s1, s2, "eager vs eager")
    P.capture()
    P.replay(); torch.cuda.synchronize(); s3 = snapshot(P)
    first_diff(s2, s3, "eager vs graph")
    P.replay(); torch.cuda.synchronize(); s4 = snapshot(P)
    first_diff(s3, s4, "graph vs graph")
    P.check_device_errors()
    print("steps:", len(P.steps), [n for _, _, n in P.steps][:12])


if __name__ == "__main__":
    main()
