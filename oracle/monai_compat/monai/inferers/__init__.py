"""monai.inferers.sliding_window_inference (0.7.0), constant blending, as called by
DosePrediction/Train/train_light_linked_model.py:152-154."""
import math

import torch
import torch.nn.functional as F


def _scan_interval(image_size, roi_size, overlap):
    out = []
    for i, r in zip(image_size, roi_size):
        if r == i:
            out.append(int(r))
        else:
            iv = int(r * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def dense_patch_slices(image_size, patch_size, scan_interval):
    nd = len(image_size)
    starts = []
    for d in range(nd):
        if scan_interval[d] == 0:
            n = 1
        else:
            n = int(math.ceil(float(image_size[d]) / scan_interval[d]))
            n = min(n for n in [n] + [k + 1 for k in range(n) if k * scan_interval[d] + patch_size[d] >= image_size[d]])
        starts.append([min(k * scan_interval[d], image_size[d] - patch_size[d]) for k in range(n)])
    grid = [[]]
    for d in range(nd):
        grid = [g + [s] for g in grid for s in starts[d]]
    return [tuple(slice(s, s + patch_size[d]) for d, s in enumerate(g)) for g in grid]


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, sw_device=None,
                             device=None, *args, **kwargs):
    if mode != "constant":
        raise NotImplementedError("compat: constant blending only")
    nd = inputs.dim() - 2
    if isinstance(roi_size, int):
        roi_size = (roi_size,) * nd
    batch = inputs.shape[0]
    orig = list(inputs.shape[2:])
    image_size = tuple(max(orig[i], roi_size[i]) for i in range(nd))
    pad = []
    for k in range(inputs.dim() - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad.extend([half, diff - half])
    inputs = F.pad(inputs, pad=pad, mode=padding_mode, value=cval)
    slices = dense_patch_slices(image_size, roi_size, _scan_interval(image_size, roi_size, overlap))
    num_win = len(slices)
    total = num_win * batch
    out_img = cnt = None
    for g in range(0, total, sw_batch_size):
        rng = range(g, min(g + sw_batch_size, total))
        idxs = [(slice(i // num_win, i // num_win + 1), slice(None)) + slices[i % num_win] for i in rng]
        prob = predictor(torch.cat([inputs[ix] for ix in idxs]), *args, **kwargs)
        if out_img is None:
            shape = [batch, prob.shape[1]] + list(image_size)
            out_img = torch.zeros(shape, dtype=prob.dtype, device=prob.device)
            cnt = torch.zeros(shape, dtype=prob.dtype, device=prob.device)
        for j, ix in zip(rng, idxs):
            out_img[ix] += prob[j - g]
            cnt[ix] += 1.0
    out_img = out_img / cnt
    final = [slice(None), slice(None)]
    for sp in range(nd):
        s = pad[(nd - 1 - sp) * 2]
        final.append(slice(s, s + orig[sp]))
    return out_img[tuple(final)]
