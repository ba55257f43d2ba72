"""Host half of the input pipeline (no GPU): NIfTI-1 reading and the RAS orientation decision."""
import numpy as np


def test_read_nifti_round_trip_and_scaling(tmp_path):
    from dose_prediction_b200 import pipeline
    from oracle import pipeline_ref
    rng = np.random.default_rng(1)
    affine = np.array([[0.0, -2.0, 0.0, 10.0], [1.5, 0.0, 0.0, -5.0], [0.0, 0.0, 3.0, 2.0], [0, 0, 0, 1.0]])
    for dtype in (np.uint8, np.int16, np.float32):
        vol = (rng.random((5, 7, 9)) * 100).astype(dtype)
        for name in ("v.nii", "v.nii.gz"):
            pipeline_ref.write_nifti(tmp_path / name, vol, affine)
            arr, aff = pipeline.read_nifti(tmp_path / name)
            assert arr.shape == (5, 7, 9) and arr.dtype == np.dtype(dtype) and np.array_equal(arr, vol)
            assert np.allclose(aff, affine)
    pipeline_ref.write_nifti(tmp_path / "s.nii", vol.astype(np.int16), affine, slope=0.5, inter=-3.0)
    arr, _ = pipeline.read_nifti(tmp_path / "s.nii")
    assert np.allclose(arr, vol.astype(np.int16) * 0.5 - 3.0)        # nibabel get_fdata semantics


def test_ras_orientation_matches_the_axis_the_affine_points_along():
    from dose_prediction_b200 import pipeline
    assert pipeline.ras_orientation(np.eye(4)) == ((0, 1, 2), (False, False, False))
    lps = np.diag([-1.0, -1.0, 1.0, 1.0])                              # ITK / DICOM LPS: flip the first two axes
    assert pipeline.ras_orientation(lps) == ((0, 1, 2), (True, True, False))
    swapped = np.array([[0.0, -2.0, 0.0, 0], [1.5, 0.0, 0.0, 0], [0.0, 0.0, 3.0, 0], [0, 0, 0, 1.0]])
    assert pipeline.ras_orientation(swapped) == ((1, 0, 2), (True, False, False))
    oblique = np.array([[0.9, 0.1, 0.0, 0], [-0.1, 0.9, 0.2, 0], [0.0, -0.2, 0.9, 0], [0, 0, 0, 1.0]])
    assert pipeline.ras_orientation(oblique) == ((0, 1, 2), (False, False, False))
