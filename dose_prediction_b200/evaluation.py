"""On-device evaluation of predicted dose volumes.

Mirrors what `Pyfer.test_step` / `LinkedNet.test_step` compute on the CPU with numpy after copying the prediction
back (DosePrediction/Train/train_light_pyfer.py:199-222, train_light_linked_model.py:171-176) through
`get_Dose_score_and_DVH_score_batch` (DosePrediction/Evaluate/evaluate_openKBP.py:149-222): post-processing,
dose score, DVH metrics of the seven OARs and three targets, and IVS over 101 isodose levels — as CUDA kernels
behind the C ABI (csrc/eval.cu), so only a few dozen floats per volume ever leave the GPU.
"""
import numpy as np
import torch

from . import _lib

OAR_NAMES = ["Brainstem", "SpinalCord", "RightParotid", "LeftParotid", "Esophagus", "Larynx", "Mandible"]
TARGET_NAMES = ["PTV70", "PTV63", "PTV56"]
STRUCTURES = OAR_NAMES + TARGET_NAMES            # order of evaluate_openKBP.py:176-186
OPENKBP_SPACING = (3.906, 3.906, 2.5)            # mm; voxels_in_tenth_of_cc = max(1, round(100 / prod(spacing)))


class DoseEvaluator:
    def __init__(self, device, n_levels=101, max_dose=70.0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dose_prediction_b200 evaluates on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.levels = torch.from_numpy(np.linspace(0, max_dose, n_levels)).to(self.device)        # float64, as the reference
        self.n_levels = n_levels
        self.acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.hist = torch.zeros(3 * (n_levels + 1), dtype=torch.int64, device=self.device)
        self.ws = torch.zeros(int(self.lib.dp_dvh_workspace_bytes()), dtype=torch.uint8, device=self.device)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def postprocess(self, prediction, possible_dose_mask, scale=70.0):
        """prediction[(mask < 1) | (prediction < 0)] = 0; prediction *= 70   (train_light_pyfer.py:210-213)."""
        p = prediction.contiguous().float()
        m = possible_dose_mask.contiguous().float()
        out = torch.empty_like(p)
        _lib.check(self.lib.dp_dose_postprocess(p.data_ptr(), m.data_ptr(), p.numel(), float(scale), out.data_ptr(),
                                                self._stream()), "dp_dose_postprocess")
        return out

    def evaluate(self, prediction, real_dose, dose_mask, structures, spacing=OPENKBP_SPACING, ivs=True):
        """One volume.  prediction / real_dose / dose_mask: [1,1,D,H,W] (or [D,H,W]) in Gy; structures: dict
        name -> mask tensor (names of STRUCTURES; missing or empty ones are skipped like the reference does).
        Returns device tensors: dose_dif [1], dvh_dif [1], dvh [n,2,5] (+ the structure names), ivs [n_levels]."""
        dev = self.device
        p = prediction.reshape(-1).contiguous().float()
        g = real_dose.to(dev).reshape(-1).contiguous().float()
        m = dose_mask.to(dev).reshape(-1).contiguous().float()
        vox = p.numel()
        s = self._stream()
        ivs_out = torch.zeros(self.n_levels, device=dev)
        dose_dif = torch.zeros(1, device=dev)
        _lib.check(self.lib.dp_dose_stats(p.data_ptr(), g.data_ptr(), m.data_ptr(), vox, self.levels.data_ptr(), self.n_levels,
                                          self.acc.data_ptr(), self.hist.data_ptr(), ivs_out.data_ptr(), dose_dif.data_ptr(), s),
                   "dp_dose_stats")
        names = []
        for n in STRUCTURES:                      # the reference stops at the first structure the batch does not carry
            if n not in structures:               # (evaluate_openKBP.py:188-189 `break`)
                break
            names.append(n)
        masks = torch.stack([structures[n].to(dev).reshape(-1).float() for n in names]).contiguous()
        is_target = torch.tensor([int(n in TARGET_NAMES) for n in names], dtype=torch.int32, device=dev)
        vt = max(1.0, float(np.round(100.0 / float(np.prod(spacing)))))
        dvh = torch.zeros((len(names), 2, 5), device=dev)
        dvh_dif = torch.zeros(1, device=dev)
        _lib.check(self.lib.dp_dvh_metrics(p.data_ptr(), g.data_ptr(), masks.data_ptr(), len(names), is_target.data_ptr(), vox,
                                           vt, self.ws.data_ptr(), dvh.data_ptr(), dvh_dif.data_ptr(), s), "dp_dvh_metrics")
        return {"dose_dif": dose_dif, "dvh_dif": dvh_dif, "dvh": dvh, "names": names, "ivs": ivs_out if ivs else None}

    @staticmethod
    def dvh_table(result):
        """{'pre<structure>_<metric>' / 'gt_<structure>_<metric>': value} like dict_DVH_dif of the reference (one D2H copy)."""
        dvh = result["dvh"].cpu().numpy()
        out = {}
        for i, n in enumerate(result["names"]):
            metrics = ["D1", "D95", "D99", "mean"] if n in TARGET_NAMES else ["D_0.1_cc", "mean"]
            cols = [0, 1, 2, 4] if n in TARGET_NAMES else [0, 4]
            if not np.any(dvh[i]):
                continue
            for mname, c in zip(metrics, cols):
                out["pre" + n + "_" + mname] = float(dvh[i, 0, c])
                out["gt_" + n + "_" + mname] = float(dvh[i, 1, c])
        return out


def dice_metric(logits, label):
    """Mean foreground Dice of one-hot(argmax(logits)) vs the label map, as Transeg.validation_step computes it with
    monai's DiceMetric(include_background=False) (OARSegmentation/train_light_transeg.py:199-216).
    logits [N,C,D,H,W] fp32 CUDA, label [N,1,D,H,W] class indices.  Returns (mean_dice [1], dice [N,C]) on the device."""
    if not logits.is_cuda:
        raise RuntimeError("dose_prediction_b200 evaluates on CUDA devices only (no CPU fallback)")
    lib = _lib.lib()
    x = logits.contiguous().float()
    lab = label.to(x.device).contiguous().float()
    N, C = x.shape[0], x.shape[1]
    vox = x.numel() // (N * C)
    counts = torch.zeros(N * 48, dtype=torch.int64, device=x.device)
    dice = torch.zeros((N, C), device=x.device)
    mean = torch.zeros(1, device=x.device)
    _lib.check(lib.dp_dice_metric(x.data_ptr(), lab.data_ptr(), N, C, vox, counts.data_ptr(), dice.data_ptr(), mean.data_ptr(),
                                  torch.cuda.current_stream(x.device).cuda_stream), "dp_dice_metric")
    return mean, dice


def hd95_metric(logits, label, percentile=95.0):
    """HausdorffDistanceMetric(include_background=False, percentile=95, reduction="mean") of one-hot(argmax(logits)) vs the label
    map, as Transeg.validation_step / test_step compute it (OARSegmentation/train_light_transeg.py:158-166,199-216).
    logits [N,C,D,H,W] fp32 CUDA, label [N,1,D,H,W] class indices.  Returns (mean [1], hd [N,C]) on the device; the mean
    follows monai's do_metric_reduction(.., "mean"): NaN entries ignored, classes first, then samples."""
    if not logits.is_cuda:
        raise RuntimeError("dose_prediction_b200 evaluates on CUDA devices only (no CPU fallback)")
    lib = _lib.lib()
    x = logits.contiguous().float()
    lab = label.to(x.device).contiguous().float()
    N, C, D, H, W = x.shape
    ws = torch.empty(int(lib.dp_hd95_workspace_bytes(C, D, H, W)), dtype=torch.uint8, device=x.device)
    hd = torch.empty((N, C), device=x.device)
    s = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        for n in range(N):
            _lib.check(lib.dp_hd95(x[n].data_ptr(), lab[n].data_ptr(), C, D, H, W, float(percentile), ws.data_ptr(),
                                   hd[n].data_ptr(), s), "dp_hd95")
    f = hd[:, 1:]
    ok = ~torch.isnan(f)
    per = torch.where(ok, f, torch.zeros_like(f)).sum(1) / ok.sum(1).clamp_min(1)
    has = ok.any(1)
    mean = (per * has).sum() / has.sum().clamp_min(1)
    return mean.reshape(1), hd
