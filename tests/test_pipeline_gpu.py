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


def test_read_nifti_orientation_and_pos_neg_crop(tmp_path):
    """LoadImaged -> Orientationd(RAS) -> RandCropByPosNegLabeld on the device vs the numpy restatements of nibabel /
    monai 0.7.0 (oracle/pipeline_ref.py), same numpy RandomState seed."""
    import numpy as np
    from dose_prediction_b200 import pipeline
    from oracle import pipeline_ref
    rng = np.random.default_rng(5)
    ct = (rng.standard_normal((20, 24, 28)) * 400).astype(np.int16)
    # voxel axes (i, j, k) -> world (A, -R, S) with anisotropic spacing: RAS needs a transpose and a flip
    affine = np.array([[0.0, -2.0, 0.0, 10.0], [1.5, 0.0, 0.0, -5.0], [0.0, 0.0, 3.0, 2.0], [0, 0, 0, 1.0]])
    for name in ("ct.nii", "ct.nii.gz"):
        pipeline_ref.write_nifti(tmp_path / name, ct, affine)
        arr, aff = pipeline.read_nifti(tmp_path / name)
        assert arr.dtype == np.int16 and np.array_equal(arr, ct) and np.allclose(aff, affine)
    perm, flips = pipeline.ras_orientation(affine)
    assert perm == (1, 0, 2) and flips == (True, False, False)
    assert pipeline.ras_orientation(np.eye(4)) == ((0, 1, 2), (False, False, False))
    pipe = pipeline.InputPipeline("cuda:0")
    x = torch.from_numpy(rng.standard_normal((3, 20, 24, 28)).astype(np.float32)).cuda()
    got = pipe.orient_ras(x, affine).cpu().numpy()
    assert np.array_equal(got, pipeline_ref.apply_orientation(x.cpu().numpy(), perm, flips))
    # RandCropByPosNegLabeld: label = 2 channels (like 'GT'), image = 9 channels (like 'Input'), 4 samples of 16^3 from 40x36x44
    S = (40, 36, 44)
    label = np.zeros((2,) + S, np.float32)
    label[0, 10:22, 8:30, 12:40] = rng.random((12, 22, 28)).astype(np.float32)
    label[1, 5:35, 4:32, 6:42] = 1.0
    image = rng.standard_normal((9,) + S).astype(np.float32)
    image[:, :3] = -1.0                                          # a slab that is neither foreground nor background
    want, want_starts = pipeline_ref.rand_crop_by_pos_neg_label([image, label], label, image, 16, 2, 1, 6, 0.0,
                                                                np.random.RandomState(123))
    outs, roi = pipe.rand_crop_by_pos_neg_label([torch.from_numpy(image).cuda(), torch.from_numpy(label).cuda()],
                                                torch.from_numpy(label).cuda(), torch.from_numpy(image).cuda(), 16, pos=2, neg=1,
                                                num_samples=6, image_threshold=0.0, rand_state=np.random.RandomState(123))
    torch.cuda.synchronize()
    assert np.array_equal(roi.cpu().numpy(), want_starts)
    assert np.array_equal(outs[0].cpu().numpy(), want[0]) and np.array_equal(outs[1].cpu().numpy(), want[1])
    # centres near the border are clamped so the crop fits (correct_crop_centers)
    assert (want_starts >= 0).all() and (want_starts + 16 <= np.asarray(S)).all()
