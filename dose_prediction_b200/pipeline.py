"""Input pipeline on the device.

Mirrors the per-volume transforms of `prepare_data` (DosePrediction/DataLoader/dataloader_OpenKBP_monai.py:160-243)
that follow file loading: axis transpose, PTV merge, CT window / scaling, dose scaling, channel stacking into
`Input` [9,...] / `GT` [2,...], and the training augmentations (intensity shift, flips, 90-degree rotations) — as two
CUDA kernels behind the C ABI (csrc/pipeline.cu).  Reading NIfTI files stays on the host (SimpleITK / nibabel are the
reference's job); the random draws are the caller's (pass the outcomes in).
"""
import torch

from . import _lib

OAR_NAMES = ["Brainstem", "SpinalCord", "RightParotid", "LeftParotid", "Esophagus", "Larynx", "Mandible"]
PTV_NAMES = ["PTV70", "PTV63", "PTV56"]


class InputPipeline:
    def __init__(self, device, a_min=-1024.0, a_max=1500.0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dose_prediction_b200 prepares inputs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.a_min, self.a_max = float(a_min), float(a_max)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def prepare(self, raw, ct_shift=0.0):
        """raw: {name: array as read from the file, shape [A,B,C]} with uint8 masks under the OAR / PTV names and
        'dose_mask', 'CT' (int16 or float32 HU) and optionally 'dose' (float32 Gy); missing structures = empty
        (Empty2FullOAR, :84-95).  Returns (Input [9,C,B,A], GT [2,C,B,A]) fp32 on the device."""
        dev = self.device
        t = {k: torch.as_tensor(v).to(dev).contiguous() for k, v in raw.items()}
        ct = t["CT"]
        A, B, C = ct.shape
        for k in OAR_NAMES + PTV_NAMES + ["dose_mask"]:
            if k in t and t[k].dtype != torch.uint8:
                t[k] = t[k].to(torch.uint8)
        ptv = _lib.ptr_array([t[n].data_ptr() if n in t else 0 for n in PTV_NAMES])
        oar = _lib.ptr_array([t[n].data_ptr() if n in t else 0 for n in OAR_NAMES])
        ct16 = ct.data_ptr() if ct.dtype == torch.int16 else None
        ct32 = None
        if ct16 is None:
            ct = ct.float().contiguous()
            ct32 = ct.data_ptr()
        dose = t["dose"].float().contiguous() if "dose" in t else None
        inp = torch.empty((9, C, B, A), device=dev)
        gt = torch.empty((2, C, B, A), device=dev)
        _lib.check(self.lib.dp_prepare_input(ptv, oar, ct16, ct32, dose.data_ptr() if dose is not None else None,
                                             t["dose_mask"].data_ptr() if "dose_mask" in t else None, A, B, C, self.a_min,
                                             self.a_max, float(ct_shift), inp.data_ptr(), gt.data_ptr(), self._stream()),
                   "dp_prepare_input")
        self._keep = (t, ct, dose)            # inputs stay alive until the stream has consumed them
        return inp, gt

    def augment(self, x, flips=(False, False, False), k=0):
        """RandFlipd on the chosen spatial axes, then RandRotate90d with k quarter turns in the (0,1) plane, on a
        [C,S0,S1,S2] fp32 tensor (apply the same arguments to Input and GT, as the reference's keyed transforms do)."""
        x = x.contiguous()
        Cc, S0, S1, S2 = x.shape
        out = torch.empty((Cc, S1, S0, S2) if (k & 1) else (Cc, S0, S1, S2), device=x.device)
        _lib.check(self.lib.dp_flip_rot90(x.data_ptr(), out.data_ptr(), Cc, S0, S1, S2, int(flips[0]), int(flips[1]),
                                          int(flips[2]), int(k), self._stream()), "dp_flip_rot90")
        return out


# ----------------------------------------------------------------------------------------------------------------------
# Round 2: the remaining pieces of prepare_data — file reading, Orientationd, RandCropByPosNegLabeld.
# monai 0.7.0 / nibabel are un-vendored dependencies (absent offline): restated from their published sources, parity
# unpinned; tests compare with the numpy restatements in oracle/pipeline_ref.py.

_NIFTI_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}


def read_nifti(path):
    """LoadImaged for the OpenKBP `.nii` / `.nii.gz` files (dataloader_OpenKBP_monai.py:166; monai's NibabelReader): a
    single-file NIfTI-1 image -> (array in the file's (i, j, k) index order, scl_slope / scl_inter applied like
    nibabel's get_fdata, 4x4 affine: sform if sform_code > 0, else the qform, else the pixdim scaling)."""
    import gzip
    import struct

    import numpy as np
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    endian = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(endian + "i", raw[:4])[0] != 348:
        raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:348] not in (b"n+1\0", b"ni1\0"):
        raise ValueError(f"{path}: bad NIfTI-1 magic {raw[344:348]!r}")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, = struct.unpack(endian + "h", raw[70:72])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(endian + "3f", raw[108:120])
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(endian + "6f", raw[256:280])
    srow = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: NIfTI datatype {datatype} not supported")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=np.dtype(endian + _NIFTI_DTYPES[datatype]), count=n, offset=int(vox_offset))
    data = data.reshape(shape, order="F")
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope) and np.isfinite(inter):
            data = data.astype(np.float64) * slope + inter
    affine = np.eye(4)
    if sform_code > 0:
        affine[:3] = srow
    elif qform_code > 0:
        qa = np.sqrt(max(0.0, 1.0 - (qb * qb + qc * qc + qd * qd)))
        R = np.array([[qa * qa + qb * qb - qc * qc - qd * qd, 2 * (qb * qc - qa * qd), 2 * (qb * qd + qa * qc)],
                      [2 * (qb * qc + qa * qd), qa * qa + qc * qc - qb * qb - qd * qd, 2 * (qc * qd - qa * qb)],
                      [2 * (qb * qd - qa * qc), 2 * (qc * qd + qa * qb), qa * qa + qd * qd - qb * qb - qc * qc]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
        affine[:3, 3] = (qx, qy, qz)
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return np.ascontiguousarray(data), affine


def ras_orientation(affine):
    """nibabel.orientations.io_orientation(affine) composed with ornt_transform(.., RAS), as monai 0.7.0 Orientation does for
    axcodes='RAS': returns (perm, flips) — output axis a reads input axis perm[a], reversed when flips[a]."""
    import numpy as np
    RZS = np.asarray(affine, dtype=np.float64)[:3, :3]
    zooms = np.sqrt(np.sum(RZS * RZS, axis=0))
    zooms[zooms == 0] = 1
    RS = RZS / zooms
    P, S, Qs = np.linalg.svd(RS, full_matrices=False)
    tol = S.max() * max(RS.shape) * np.finfo(S.dtype).eps
    keep = S > tol
    R = np.dot(P[:, keep], Qs[keep])
    perm, flips = [None] * 3, [False] * 3
    for in_ax in range(3):
        col = R[:, in_ax]
        if not np.allclose(col, 0):
            out_ax = int(np.argmax(np.abs(col)))
            perm[out_ax] = in_ax
            flips[out_ax] = bool(col[out_ax] < 0)
            R[out_ax, :] = 0
    if any(p is None for p in perm):
        raise ValueError("affine does not map every voxel axis to a world axis")
    return tuple(perm), tuple(flips)


def _pinned_ints(vals):
    return torch.tensor(vals, dtype=torch.int32)


def _orient(self, x, affine):
    """Orientationd(axcodes='RAS') on a [C,S0,S1,S2] fp32 device tensor."""
    perm, flips = ras_orientation(affine)
    x = x.contiguous()
    C, S = x.shape[0], x.shape[1:]
    out = torch.empty((C, S[perm[0]], S[perm[1]], S[perm[2]]), device=x.device)
    _lib.check(self.lib.dp_permute_flip(x.data_ptr(), out.data_ptr(), C, S[0], S[1], S[2], perm[0], perm[1], perm[2],
                                        int(flips[0]), int(flips[1]), int(flips[2]), self._stream()), "dp_permute_flip")
    return out


def _rand_crop_by_pos_neg_label(self, tensors, label, image, spatial_size, pos=1.0, neg=1.0, num_samples=1, image_threshold=0.0,
                                rand_state=None):
    """RandCropByPosNegLabeld (dataloader_OpenKBP_monai.py:206-215; provided_dataset.py:158-167): `tensors` = list of
    [C_i,S0,S1,S2] fp32 device tensors cropped alike; label [Cl,...] (foreground = any channel > 0), image [Ci,...] or None.
    rand_state: numpy RandomState (monai's self.R) — the draws follow generate_pos_neg_label_crop_centers exactly:
    rand() < pos_ratio picks foreground, randint(len(indices)) the voxel.  Returns ([num_samples, C_i, R, R, R] per tensor,
    crop origins as an int32 device tensor)."""
    import numpy as np
    R = int(spatial_size)
    rs = rand_state if rand_state is not None else np.random.RandomState()
    label = label.contiguous().float()
    S0, S1, S2 = label.shape[1:]
    vox = S0 * S1 * S2
    img = image.contiguous().float() if image is not None else None
    nblk = (vox + 4095) // 4096
    counts = torch.empty((nblk, 2), dtype=torch.int32, device=self.device)
    _lib.check(self.lib.dp_posneg_count(label.data_ptr(), label.shape[0], img.data_ptr() if img is not None else None,
                                        img.shape[0] if img is not None else 0, float(image_threshold), vox, counts.data_ptr(),
                                        self._stream()), "dp_posneg_count")
    c = counts.cpu().numpy().astype(np.int64)                 # the one host round trip: numpy's randint needs the totals
    cum = np.cumsum(c, axis=0)
    n_fg, n_bg = int(cum[-1, 0]), int(cum[-1, 1])
    pos_ratio = pos / (pos + neg)
    if n_fg == 0 or n_bg == 0:                                # monai: warn and use whichever set exists
        if n_fg == 0 and n_bg == 0:
            raise ValueError("No sampling location available.")
        pos_ratio = 0 if n_fg == 0 else 1
    picks = []
    for _ in range(int(num_samples)):
        want_fg = rs.rand() < pos_ratio
        total = n_fg if want_fg else n_bg
        k = int(rs.randint(total))
        col = 0 if want_fg else 1
        blk = int(np.searchsorted(cum[:, col], k, side="right"))
        rank = k - (int(cum[blk - 1, col]) if blk > 0 else 0)
        picks += [blk, rank, int(want_fg)]
    pick = _pinned_ints(picks).to(self.device)
    roi = torch.empty((int(num_samples), 3), dtype=torch.int32, device=self.device)
    srcs = [t.contiguous().float() for t in tensors]
    outs = [torch.empty((int(num_samples), t.shape[0], R, R, R), device=self.device) for t in srcs]
    _lib.check(self.lib.dp_posneg_crop(label.data_ptr(), label.shape[0], img.data_ptr() if img is not None else None,
                                       img.shape[0] if img is not None else 0, float(image_threshold), S0, S1, S2, R,
                                       int(num_samples), pick.data_ptr(), roi.data_ptr(), len(srcs),
                                       _lib.ptr_array([t.data_ptr() for t in srcs]), _lib.int_array([t.shape[0] for t in srcs]),
                                       _lib.ptr_array([o.data_ptr() for o in outs]), self._stream()), "dp_posneg_crop")
    self._keep_crop = (label, img, srcs, pick)
    return outs, roi


InputPipeline.orient_ras = _orient
InputPipeline.rand_crop_by_pos_neg_label = _rand_crop_by_pos_neg_label
