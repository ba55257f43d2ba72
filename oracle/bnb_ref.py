"""TEST INFRASTRUCTURE — numpy restatement of bitsandbytes 0.40.2's block-wise 8-bit state quantisation
(`functional.quantize_blockwise` / `dequantize_blockwise` with blocksize 2048, as the 8-bit optimizers use it).

bitsandbytes is an un-vendored dependency of the reference (requirements.txt: bitsandbytes==0.40.2; call site
DosePrediction/Train/train_light_pyfer.py:194-197) and is absent offline: PARITY UNPINNED — this restates the published
algorithm (per-block absmax, nearest entry of the sorted 256-value code book) and is what tests/ compare the CUDA kernels to.
"""
import numpy as np

BLOCK = 2048


def quantize_blockwise(x, qmap):
    x = np.asarray(x, dtype=np.float32).reshape(-1)
    q = np.asarray(qmap, dtype=np.float32)
    n = x.size
    nb = (n + BLOCK - 1) // BLOCK
    pad = np.zeros(nb * BLOCK, dtype=np.float32)
    pad[:n] = x
    blocks = pad.reshape(nb, BLOCK)
    absmax = np.abs(blocks).max(axis=1)
    inv = np.where(absmax > 0, 1.0 / np.maximum(absmax, 1e-45), 0.0).astype(np.float32)
    t = blocks * inv[:, None]
    hi = np.clip(np.searchsorted(q, t, side="right"), 1, 255)       # q[hi-1] <= t < q[hi]
    lo = hi - 1
    codes = np.where(t - q[lo] <= q[hi] - t, lo, hi).astype(np.uint8)
    return codes.reshape(-1)[:n], absmax.astype(np.float32)


def dequantize_blockwise(codes, absmax, qmap):
    codes = np.asarray(codes, dtype=np.uint8).reshape(-1)
    q = np.asarray(qmap, dtype=np.float32)
    scale = np.repeat(np.asarray(absmax, dtype=np.float32), BLOCK)[:codes.size]
    return q[codes] * scale
