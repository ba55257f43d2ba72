"""TEST INFRASTRUCTURE — numpy restatement of the reference's per-volume evaluation
(DosePrediction/Evaluate/evaluate_openKBP.py).  Pinned against the reference file itself by tests/test_oracle.py
(imported through oracle.ref_loader with stand-ins for its unused SimpleITK / matplotlib imports)."""
import numpy as np

OARS = ["Brainstem", "SpinalCord", "RightParotid", "LeftParotid", "Esophagus", "Larynx", "Mandible"]
TARGETS = ["PTV70", "PTV63", "PTV56"]


def postprocess(prediction, possible_dose_mask):
    """train_light_pyfer.py:210-213."""
    p = np.array(prediction, dtype=np.float32, copy=True)
    p[np.logical_or(possible_dose_mask < 1, p < 0)] = 0
    return 70.0 * p


def dose_dif(pred, gt, possible_dose_mask):
    """evaluate_openKBP.py:42-48 get_3D_Dose_dif."""
    sel = possible_dose_mask > 0
    return float(np.mean(np.abs(pred[sel] - gt[sel])))


def ivs(pred, gt, level):
    """evaluate_openKBP.py:17-39 IVS (possible_dose_mask=None as called at :166)."""
    a, b = pred >= level, gt >= level
    return 2 * np.sum(a * b) / (np.sum(a) + np.sum(b))


def dvh_metrics(dose, mask, mode, spacing):
    """evaluate_openKBP.py:51-81 get_DVH_metrics."""
    roi = dose[mask > 0]
    if mode == "target":
        return {"D1": np.percentile(roi, 99), "D95": np.percentile(roi, 5), "D99": np.percentile(roi, 1), "mean": np.mean(roi)}
    vt = np.maximum(1, np.round(100 / np.prod(spacing)))
    return {"D_0.1_cc": np.percentile(roi, 100 - vt / len(roi) * 100), "mean": np.mean(roi)}


def evaluate(pred, gt, possible_dose_mask, structures, spacing):
    """evaluate_openKBP.py:149-222 get_Dose_score_and_DVH_score_batch for one volume (arrays [D,H,W])."""
    out = {"dose_dif": dose_dif(pred, gt, possible_dose_mask),
           "ivs": [ivs(pred, gt, lv) for lv in np.linspace(0, 70, 101)], "table": {}}
    difs = []
    for name in OARS + TARGETS:
        if name not in structures:
            break
        m = structures[name]
        if np.any(m):
            mode = "target" if name in TARGETS else "OAR"
            a, b = dvh_metrics(pred, m, mode, spacing), dvh_metrics(gt, m, mode, spacing)
            for k in b:
                difs.append(abs(b[k] - a[k]))
                out["table"]["pre" + name + "_" + k] = float(a[k])
                out["table"]["gt_" + name + "_" + k] = float(b[k])
    out["dvh_dif"] = float(np.mean(difs))
    return out


def dice_metric(logits, label):
    """monai 0.7.0 DiceMetric(include_background=False, reduction="mean", get_not_nans=False) applied to
    post_pred = AsDiscrete(argmax=True, to_onehot=True) and post_label (OARSegmentation/config.py:69-70,
    train_light_transeg.py:199-216): compute_meandice gives 2|y & y_pred| / (|y| + |y_pred|) per (batch, class), NaN when
    the class is absent from y; monai.metrics.utils.do_metric_reduction(f, "mean") then averages in TWO steps, ignoring
    NaNs: over the classes of each sample first, then over the samples that have at least one class present (0 when no
    sample has one).  (un-vendored monai code, restated.)"""
    n_cls = logits.shape[1]
    pred = logits.argmax(1)
    lab = label[:, 0].astype(np.int64)
    per_sample = []
    for b in range(logits.shape[0]):
        vals = []
        for c in range(1, n_cls):
            y, yp = lab[b] == c, pred[b] == c
            if y.sum() > 0:
                vals.append(2.0 * np.logical_and(y, yp).sum() / (y.sum() + yp.sum()))
        if vals:
            per_sample.append(float(np.mean(vals)))
    return float(np.mean(per_sample)) if per_sample else 0.0


def hd95_metric(logits, label, percentile=95):
    """monai 0.7.0 HausdorffDistanceMetric(include_background=False, percentile=95) on AsDiscrete(argmax, to_onehot) vs the
    one-hot label (OARSegmentation/train_light_transeg.py:158-166,199-216; un-vendored monai code, restated with the same scipy
    calls it makes): get_mask_edges (crop to the joint bounding box, mask ^ binary_erosion(mask)), get_surface_distance
    (distance_transform_edt of the other surface's complement), np.percentile both ways, max.  Returns hd [N, C-1]."""
    from scipy.ndimage import binary_erosion, distance_transform_edt
    n_cls = logits.shape[1]
    pred = logits.argmax(1)
    lab = label[:, 0].astype(np.int64)
    out = np.empty((logits.shape[0], n_cls - 1))

    def one_way(e_src, e_dst):
        if not np.any(e_dst):
            dis = np.inf * np.ones_like(e_dst, dtype=np.float64)
        else:
            if not np.any(e_src):
                dis = np.inf * np.ones_like(e_dst, dtype=np.float64)
                return np.asarray(dis[e_dst])
            dis = distance_transform_edt(~e_dst)
        return np.asarray(dis[e_src])

    for b in range(logits.shape[0]):
        for c in range(1, n_cls):
            sp, sg = pred[b] == c, lab[b] == c
            if not np.any(sp | sg):
                out[b, c - 1] = np.nan
                continue
            idx = np.nonzero(sp | sg)
            box = tuple(slice(i.min(), i.max() + 1) for i in idx)
            sp, sg = sp[box], sg[box]
            ep, eg = binary_erosion(sp) ^ sp, binary_erosion(sg) ^ sg
            d = []
            for src, dst in ((ep, eg), (eg, ep)):
                sd = one_way(src, dst)
                d.append(np.nan if sd.shape == (0,) else np.percentile(sd, percentile))
            out[b, c - 1] = max(d[0], d[1]) if not (np.isnan(d[0]) or np.isnan(d[1])) else np.nan
    return out
