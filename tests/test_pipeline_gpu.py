"""Input pipeline kernels vs the numpy restatement of the reference's transforms: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _raw(shape, seed, ct_dtype):
    rng = np.random.default_rng(seed)
    raw = {"CT": rng.integers(-2000, 3000, shape).astype(ct_dtype),
           "dose": (rng.random(shape) * 75).astype(np.float32),
           "dose_mask": (rng.random(shape) < 0.6).astype(np.uint8)}
    for n in ("Brainstem", "SpinalCord", "LeftParotid", "Mandible", "PTV70", "PTV56"):     # some structures absent
        raw[n] = (rng.random(shape) < 0.1).astype(np.uint8)
    return raw


@pytest.mark.parametrize("shape,ct_dtype,shift", [((40, 24, 72), np.int16, 0.0), ((33, 17, 50), np.float32, 0.0625)])
def test_prepare_matches_reference_transforms(shape, ct_dtype, shift):
    from dose_prediction_b200.pipeline import InputPipeline
    from oracle import pipeline_ref
    raw = _raw(shape, 3, ct_dtype)
    pipe = InputPipeline(DEV)
    inp, gt = pipe.prepare(raw, ct_shift=shift)
    torch.cuda.synchronize()
    want_inp, want_gt = pipeline_ref.prepare(raw, ct_shift=shift)
    assert inp.shape == want_inp.shape and gt.shape == want_gt.shape
    assert np.array_equal(inp.cpu().numpy(), want_inp)
    assert np.array_equal(gt.cpu().numpy(), want_gt)


@pytest.mark.parametrize("flips,k", [((True, False, False), 0), ((False, True, True), 1), ((True, True, False), 2),
                                     ((False, False, False), 3), ((True, True, True), 3)])
def test_flip_rot90_matches_numpy(flips, k):
    from dose_prediction_b200.pipeline import InputPipeline
    from oracle import pipeline_ref
    x = np.random.default_rng(5).random((3, 10, 14, 9)).astype(np.float32)
    out = InputPipeline(DEV).augment(torch.from_numpy(x).to(DEV), flips, k)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), pipeline_ref.augment(x, flips, k))
