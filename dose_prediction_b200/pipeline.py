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
