"""On-device evaluation vs the oracle (oracle/eval_ref.py) and the fixture made from the reference's
own evaluate_openKBP.py functions (tests/golden/eval64.npz).  Integer work (ROI sizes, histograms, order statistics)
is exact; the reported floats agree to fp32 rounding (tolerances below)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(size, seed):
    from dose_prediction_b200 import synth
    v = synth.make_volume(size, seed=seed)
    gt_gy = (v["gt"][0, 0].numpy() * 70).astype(np.float32)
    pmask = v["gt"][0, 1].numpy()
    raw = (gt_gy / 70 + np.random.default_rng(7).normal(0, 0.03, gt_gy.shape)).astype(np.float32)
    return v, raw, gt_gy, pmask


def test_evaluator_matches_reference_fixture_and_oracle():
    from dose_prediction_b200 import synth
    from dose_prediction_b200.evaluation import DoseEvaluator
    from oracle import eval_ref
    v, raw, gt_gy, pmask = _inputs(64, 99)
    ev = DoseEvaluator(DEV)
    pred = ev.postprocess(torch.from_numpy(raw).to(DEV), torch.from_numpy(pmask).to(DEV))
    want_pred = eval_ref.postprocess(raw, pmask)
    assert np.array_equal(pred.cpu().numpy(), want_pred)                      # bit-exact post-processing
    st = synth.structures(v)
    res = ev.evaluate(pred, torch.from_numpy(gt_gy), torch.from_numpy(pmask), st)
    torch.cuda.synchronize()
    g = np.load(os.path.join(GOLDEN, "eval64.npz"))
    assert abs(float(res["dose_dif"]) - float(g["dose_dif"])) <= 2e-6 * float(g["dose_dif"])
    assert np.allclose(res["ivs"].cpu().numpy(), g["ivs"], rtol=0, atol=1e-6)
    table = ev.dvh_table(res)
    keys = [k for k in g.files if k.startswith("pre") or k.startswith("gt_")]
    assert sorted(keys) == sorted(table)
    for k in keys:
        assert abs(table[k] - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), k
    assert abs(float(res["dvh_dif"]) - float(g["dvh_dif"])) <= 1e-5


def test_evaluator_edge_cases():
    """empty structures are skipped, tiny ROIs (n = 2, 4) and negative / zero doses select the right order statistics."""
    from dose_prediction_b200.evaluation import DoseEvaluator
    from oracle import eval_ref
    rng = np.random.default_rng(3)
    S = 24
    pred = rng.normal(20, 15, (S, S, S)).astype(np.float32)
    gt = rng.normal(20, 15, (S, S, S)).astype(np.float32)
    pred[rng.random(pred.shape) < 0.2] = 0.0
    pmask = (rng.random(pred.shape) < 0.7).astype(np.float32)
    st = {"Brainstem": np.zeros((S, S, S), np.float32), "SpinalCord": np.zeros((S, S, S), np.float32),
          "RightParotid": (rng.random(pred.shape) < 0.1).astype(np.float32), "PTV70": np.zeros((S, S, S), np.float32),
          "PTV63": (rng.random(pred.shape) < 0.3).astype(np.float32)}
    for n in ("LeftParotid", "Esophagus", "Larynx", "Mandible", "PTV56"):
        st[n] = np.zeros((S, S, S), np.float32)
    st["Mandible"][5:9, 5:9, 5:9] = 1.0
    st["SpinalCord"][3, 4, 5:9] = 1.0                    # n = 4 (an OAR below 3 voxels makes the reference's np.percentile raise)
    st["PTV70"][1, 1, 1] = st["PTV70"][2, 2, 2] = 1.0    # n = 2
    ev = DoseEvaluator(DEV)
    res = ev.evaluate(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt), torch.from_numpy(pmask),
                      {k: torch.from_numpy(m) for k, m in st.items()})
    torch.cuda.synchronize()
    want = eval_ref.evaluate(pred, gt, pmask, st, (3.906, 3.906, 2.5))
    table = ev.dvh_table(res)
    assert sorted(table) == sorted(want["table"])
    for k, val in want["table"].items():
        assert abs(table[k] - val) <= 2e-6 * max(1.0, abs(val)), k
    assert abs(float(res["dose_dif"]) - want["dose_dif"]) <= 2e-6 * want["dose_dif"]
    assert np.allclose(res["ivs"].cpu().numpy(), np.array(want["ivs"]), rtol=0, atol=1e-6)
    assert abs(float(res["dvh_dif"]) - want["dvh_dif"]) <= 1e-5


def test_dice_metric_matches_oracle():
    from dose_prediction_b200 import synth
    from dose_prediction_b200.evaluation import dice_metric
    from oracle import eval_ref
    vol = synth.make_batch(2, 32, seed=8)
    label = synth.oar_labels(vol["oars"])
    label[1][label[1] == 3] = 0                                   # class 3 absent from the second volume
    torch.manual_seed(0)
    onehot = torch.nn.functional.one_hot(label[:, 0].long(), 8).permute(0, 4, 1, 2, 3).float()
    logits = 2.0 * onehot + torch.randn(2, 8, 32, 32, 32)
    mean, dice = dice_metric(logits.to(DEV), label.to(DEV))
    torch.cuda.synchronize()
    want = eval_ref.dice_metric(logits.numpy(), label.numpy())
    assert abs(float(mean) - want) < 1e-6
    assert torch.isnan(dice[1, 3]).item() and not torch.isnan(dice[0, 3]).item()


def test_hd95_matches_the_scipy_restatement_of_monai():
    """dp_hd95 (brute-force exact nearest-surface search + d^2 histogram) vs oracle/eval_ref.hd95_metric, which makes the same
    scipy calls monai 0.7.0's HausdorffDistanceMetric makes (binary_erosion, distance_transform_edt, np.percentile)."""
    import numpy as np
    from dose_prediction_b200 import synth
    from dose_prediction_b200.evaluation import hd95_metric
    from oracle import eval_ref
    vol = synth.make_batch(2, 48, seed=8)
    label = synth.oar_labels(vol["oars"])
    label[1][label[1] == 3] = 0                                   # class 3 absent from the second volume's label
    torch.manual_seed(0)
    onehot = torch.nn.functional.one_hot(label[:, 0].long(), 8).permute(0, 4, 1, 2, 3).float()
    logits = 1.5 * onehot + torch.randn(2, 8, 48, 48, 48)          # noisy prediction: ragged surfaces, stray voxels
    logits[0, 5] = -10.0                                           # class 5 never predicted in the first volume
    mean, hd = hd95_metric(logits.to(DEV), label.to(DEV))
    torch.cuda.synchronize()
    want = eval_ref.hd95_metric(logits.numpy(), label.numpy())
    got = hd[:, 1:].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    assert fin.sum() >= 10 and np.allclose(got[fin], want[fin], rtol=1e-6, atol=1e-5)
    assert np.isfinite(float(mean)) or np.isinf(float(mean))
