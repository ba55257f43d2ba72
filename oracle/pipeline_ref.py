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


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: LoadImaged / Orientationd / RandCropByPosNegLabeld.  monai 0.7.0 and nibabel are un-vendored and absent offline:
# restated from their published sources (parity unpinned); the functions below are what tests/ compare the device path to.
def write_nifti(path, data, affine, slope=1.0, inter=0.0):
    """minimal single-file NIfTI-1 writer (test fixture generator for pipeline.read_nifti; sform carries the affine)."""
    import gzip
    import struct
    data = np.asarray(data)
    code = {np.dtype("u1"): 2, np.dtype("i2"): 4, np.dtype("i4"): 8, np.dtype("f4"): 16, np.dtype("f8"): 64}[data.dtype]
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, code)
    struct.pack_into("<h", hdr, 72, data.dtype.itemsize * 8)
    zooms = np.sqrt((np.asarray(affine)[:3, :3] ** 2).sum(0))
    struct.pack_into("<8f", hdr, 76, 1.0, *zooms, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<3f", hdr, 108, 352.0, slope, inter)
    struct.pack_into("<2h", hdr, 252, 0, 1)                        # qform_code 0, sform_code 1
    struct.pack_into("<12f", hdr, 280, *np.asarray(affine, dtype=np.float64)[:3].reshape(-1))
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + data.tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


def apply_orientation(arr, perm, flips):
    """nibabel.orientations.apply_orientation on the spatial axes of [C,S0,S1,S2]: output axis a = input axis perm[a], flipped."""
    out = np.transpose(arr, (0,) + tuple(p + 1 for p in perm))
    for a, f in enumerate(flips):
        if f:
            out = np.flip(out, axis=a + 1)
    return np.ascontiguousarray(out)


def rand_crop_by_pos_neg_label(arrays, label, image, spatial_size, pos, neg, num_samples, image_threshold, rand_state):
    """monai 0.7.0 RandCropByPosNegLabeld: map_binary_to_indices + generate_pos_neg_label_crop_centers +
    correct_crop_centers + SpatialCrop(roi_center, roi_size)."""
    R = int(spatial_size)
    shape = label.shape[1:]
    label_flat = np.any(label, axis=0).ravel()
    fg = np.nonzero(label_flat)[0]
    if image is not None:
        img_flat = np.any(image > image_threshold, axis=0).ravel()
        bg = np.nonzero(np.logical_and(img_flat, ~label_flat))[0]
    else:
        bg = np.nonzero(~label_flat)[0]
    pos_ratio = pos / (pos + neg)
    if not len(fg) or not len(bg):
        if not len(fg) and not len(bg):
            raise ValueError("No sampling location available.")
        pos_ratio = 0 if not len(fg) else 1
    outs = [[] for _ in arrays]
    starts = []
    for _ in range(num_samples):
        idx = fg if rand_state.rand() < pos_ratio else bg
        center = list(np.unravel_index(idx[rand_state.randint(len(idx))], shape))
        valid_start = np.floor_divide([R] * 3, 2)
        valid_end = np.subtract(np.asarray(shape) + np.array(1), np.asarray([R] * 3) / np.array(2)).astype(np.uint16)
        for i in range(3):
            if valid_start[i] == valid_end[i]:
                valid_end[i] += 1
        for i, c in enumerate(center):
            ci = c
            if c < valid_start[i]:
                ci = valid_start[i]
            if c >= valid_end[i]:
                ci = valid_end[i] - 1
            center[i] = int(ci)
        start = [max(c - R // 2, 0) for c in center]
        starts.append(start)
        for o, a in zip(outs, arrays):
            o.append(a[:, start[0]:start[0] + R, start[1]:start[1] + R, start[2]:start[2] + R])
    return [np.stack(o) for o in outs], np.asarray(starts, dtype=np.int32)
