"""TEST INFRASTRUCTURE — numpy restatement of the per-volume transforms of
DosePrediction/DataLoader/dataloader_OpenKBP_monai.py (prepare_data, :160-243).  monai's dictionary transforms used
there (Transposed, AddChanneld, ConcatItemsd, RandShiftIntensityd, RandFlipd, RandRotate90d) are un-vendored
monai 0.7.0 code; they reduce to the numpy calls below (np.transpose, np.stack, +offset, np.flip, np.rot90).
Parity unpinned for those monai pieces; the reference's own classes (NormalizePTVTr, MyIntensityNormalTransform,
NormalizeDoseTr, Empty2FullOAR) are restated line by line and pinned by tests/golden/pipeline12.npz, which
oracle/make_golden.py produces by running those classes themselves."""
import numpy as np

OAR_NAMES = ["Brainstem", "SpinalCord", "RightParotid", "LeftParotid", "Esophagus", "Larynx", "Mandible"]


def prepare(raw, a_min=-1024, a_max=1500, ct_shift=0.0):
    d = dict(raw)
    mask = np.zeros(d["CT"].shape, np.uint8)
    for name in OAR_NAMES + ["PTV70", "PTV63", "PTV56"]:          # Empty2FullOAR, :84-95
        d.setdefault(name, mask.copy())
    keys = ["PTV70", "PTV63", "PTV56"] + OAR_NAMES + ["CT", "dose", "dose_mask"]
    for k in keys:                                                 # Transposed(indices=[2,1,0]), :173
        if k in d:
            d[k] = np.transpose(d[k], (2, 1, 0))
    d["PTV"] = 70.0 / 70. * d["PTV70"] + 63.0 / 70. * d["PTV63"] + 56.0 / 70. * d["PTV56"]       # NormalizePTVTr, :113-125
    ct = np.clip(d["CT"], a_min=a_min, a_max=a_max)               # MyIntensityNormalTransform, :137-146
    d["CT"] = ct.astype(np.float32) / 1000.
    if ct_shift:
        d["CT"] = (d["CT"] + np.float32(ct_shift)).astype(np.float32)     # RandShiftIntensityd, :189-193
    if "dose" in d:
        d["dose"] = d["dose"] / 70.0                               # NormalizeDoseTr, :128-134
    inp = np.stack([d["PTV"]] + [d[n] for n in OAR_NAMES] + [d["CT"]]).astype(np.float32)        # ConcatItemsd, :195-197
    gt = np.stack([d.get("dose", np.zeros_like(d["CT"])), d.get("dose_mask", mask.transpose(2, 1, 0))]).astype(np.float32)
    return inp, gt


def augment(x, flips=(False, False, False), k=0):
    """RandFlipd(spatial_axis=[a]) for the chosen axes, then RandRotate90d(k, spatial_axes=(0,1)) on [C,S0,S1,S2]."""
    for a, f in enumerate(flips):
        if f:
            x = np.flip(x, axis=a + 1)
    if k:
        x = np.rot90(x, k, axes=(1, 2))
    return np.ascontiguousarray(x)
