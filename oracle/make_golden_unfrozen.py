"""TEST INFRASTRUCTURE — tests/golden/train32_unfrozen.npz from the REAL reference modules: one `Pyfer(freeze=False)`
training step (DosePrediction/Train/train_light_pyfer.py:61-88,122-143 with GenLoss(..., freez=False), loss.py:114-115).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_unfrozen
Same weights / inputs as make_golden's train32.npz; every parameter requires grad.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dose_prediction_b200 import synth  # noqa: E402
from oracle import ref_loader, synth_ckpt, torch_ref  # noqa: E402
from oracle.make_golden import DOSE_SEED, OUT, _np  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    vol = synth.make_batch(2, 32, seed=1234)
    loss_mod = ref_loader.loss().GenLoss(im_size=32)
    tm = ref_loader.build_dose(32).train()
    tm.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(tm), DOSE_SEED), strict=True)
    params = list(tm.named_parameters())
    opt = torch.optim.AdamW([p_ for _, p_ in params], lr=1e-4, weight_decay=1e-4)
    loss = loss_mod(tm(vol["dose_input"]), vol["gt"], casecade=True, freez=False, delta1=10, delta2=8)
    loss.backward()
    rec = {"loss": np.float64(loss.item())}
    names, norms = [], []
    for n, p_ in params:
        if p_.grad is None:               # never touched by forward (cls_token, net_B.out, ...): no gradient, no update
            continue
        names.append(n)
        norms.append(float(p_.grad.double().norm()))
        rec["g/" + n] = _np(p_.grad.flatten()[torch_ref.sample_idx(p_.grad.numel())])
    opt.step()
    for n in names:
        p_ = dict(params)[n]
        rec["p/" + n] = _np(p_.detach().flatten()[torch_ref.sample_idx(p_.numel())])
    rec["names"] = np.array(names)
    rec["norms"] = np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "train32_unfrozen.npz"), **rec)
    print("train32_unfrozen.npz: loss %.6f, %d parameters with gradients" % (loss.item(), len(names)))


if __name__ == "__main__":
    main()
