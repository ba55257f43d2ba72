"""Per-launch device-time table of the cascade plan (CUDA events), written to gpurun_out/layers_<tag>.csv."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dose_prediction_b200 import synth  # noqa: E402
from dose_prediction_b200.cascade import CascadePlan  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    tag = sys.argv[3] if len(sys.argv) > 3 else f"b{B}"
    dev = torch.device("cuda:0")
    seg, dose = bench.build_models(S, dev)
    casc = CascadePlan(seg, dose, B, S, dev)
    v = synth.make_batch(B, S, seed=1234)
    casc.ct.copy_(v["ct"]); casc.ptv.copy_(v["ptv"])
    for _ in range(2):
        casc.run()
    torch.cuda.synchronize()
    rows = casc.plan.profile_launches()
    rows = casc.plan.profile_launches()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"layers_{tag}.csv"), "w") as f:
        f.write("idx,family,label,ms,gflop\n")
        for i, (n, l, ms, fl) in enumerate(rows):
            f.write(f"{i},{n},{l},{ms:.4f},{fl / 1e9:.3f}\n")
    tot = sum(r[2] for r in rows)
    print(f"total {tot:.2f} ms for batch {B}; top launches:")
    for n, l, ms, fl in sorted(rows, key=lambda r: -r[2])[:40]:
        print(f"  {ms:8.3f} ms  {n:18s} {l}")
    casc.plan.check_device_errors()


if __name__ == "__main__":
    main()
