"""TEST INFRASTRUCTURE — how sensitive is the fp32 reference dose net to flipped OAR-mask voxels?

    python -m oracle.sensitivity_probe [--size 64]

The cascade's end-to-end dose (seg argmax -> masks -> dose net) can only agree with the reference's as far as the masks
agree (>= 99.9 % of the voxels, north_star).  With the random-init weights available offline the reference dose net turns
out to be chaotic in its mask inputs; this script measures it on the oracle alone (no CUDA code involved): it flips a
fraction of the argmax voxels and reports the relative L2 change of the fp32 dose.  bench.py reports
`dose_rel_l2_end_to_end` beside the gated same-input `dose_rel_l2` for that reason.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    a = ap.parse_args()
    import bench
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    seg, dose = bench.build_models(a.size)
    ssd, dsd = seg.state_dict(), dose.state_dict()
    vol = synth.make_volume(a.size, seed=1234)
    with torch.no_grad():
        logits = torch_ref.oar_transeg_forward(ssd, vol["ct"])
        st = torch_ref.handoff(logits, vol["ptv"], vol["ct"])
        d0 = torch_ref.dose_pyfer_forward(dsd, st)[1][0]
        gen = torch.Generator().manual_seed(0)
        for frac in (1e-5, 1e-4, 1e-3):
            lg = logits.clone()
            mask = torch.rand(lg.shape[2:], generator=gen) < frac
            lg[:, :, mask] = lg[:, :, mask].roll(1, dims=1)                 # a different class wins at those voxels
            st2 = torch_ref.handoff(lg, vol["ptv"], vol["ct"])
            d1 = torch_ref.dose_pyfer_forward(dsd, st2)[1][0]
            print(f"flipped {frac:g} of the voxels: structures agree {(st2 == st).float().mean().item():.6f}, "
                  f"fp32 reference dose moves by rel-L2 {torch_ref.rel_l2(d1, d0):.4f}")


if __name__ == "__main__":
    main()
