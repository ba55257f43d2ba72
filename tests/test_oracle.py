"""Oracle pinning (CPU): oracle/torch_ref.py against the fixtures produced by the reference's own
modules (oracle/make_golden.py), and — when /root/reference is present — against the modules live."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_manifest
from dose_prediction_b200 import synth
from oracle import ref_loader, synth_ckpt, torch_ref

TOL = 2e-5   # fp32 CPU conv reduction order differs between hosts / thread counts


def _resize_manifest(man, tokens):
    out = []
    for k, shape, *_ in man:
        if k.endswith("position_embeddings"):
            shape = [1, tokens, shape[2]]
        out.append((k, shape))
    return out


@pytest.fixture(scope="module")
def dose_sd32():
    return synth_ckpt.make_state_dict(_resize_manifest(load_manifest("dose_pyfer"), 8), seed=0)


@pytest.fixture(scope="module")
def seg_sd32():
    return synth_ckpt.make_state_dict(_resize_manifest(load_manifest("oar_transeg"), 8), seed=1)


def test_dose_pyfer_matches_reference_fixture(dose_sd32):
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "dose32.npz"))
    with torch.no_grad():
        out = torch_ref.dose_pyfer_forward(dose_sd32, vol["dose_input"])
    assert torch_ref.rel_l2(out[0], torch.from_numpy(g["out_A"])) < TOL
    for i, t in enumerate(out[1]):
        assert t.shape == g[f"d{i}"].shape
        assert torch_ref.rel_l2(t, torch.from_numpy(g[f"d{i}"])) < TOL


def test_oar_transeg_matches_reference_fixture(seg_sd32):
    vol = synth.make_batch(2, 32, seed=1234)
    g = torch.from_numpy(np.load(os.path.join(GOLDEN, "seg32.npz"))["logits"])
    with torch.no_grad():
        out = torch_ref.oar_transeg_forward(seg_sd32, vol["ct"])
    assert torch_ref.rel_l2(out, g) < TOL
    assert (out.argmax(1) == g.argmax(1)).float().mean().item() > 0.9999


def test_handoff_and_cascade_match_reference_fixture(dose_sd32, seg_sd32):
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "cascade32.npz"))
    ct, ptv = vol["ct"][:1], vol["ptv"][:1]
    with torch.no_grad():
        logits = torch_ref.oar_transeg_forward(seg_sd32, ct)
        st = torch_ref.handoff(logits, ptv, ct)
        agree = (st == torch.from_numpy(g["structures"])).float().mean().item()
        assert agree > 0.9999            # argmax near-ties may flip with the host's fp32 summation order
        dose = torch_ref.dose_pyfer_forward(dose_sd32, torch.from_numpy(g["structures"]))[1][0]
    assert torch_ref.rel_l2(dose, torch.from_numpy(g["dose"])) < TOL


def test_gen_loss_matches_reference_fixture():
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "dose32.npz"))
    preds = [torch.from_numpy(g["out_A"]), [torch.from_numpy(g[f"d{i}"]) for i in range(4)]]
    want = float(np.load(os.path.join(GOLDEN, "genloss32.npz"))["loss"])
    got = float(torch_ref.gen_loss(preds, vol["gt"]))
    assert abs(got - want) <= 1e-5 * abs(want)


def test_train_step_matches_reference_fixture(dose_sd32):
    """oracle train step (train-mode BN, GenLoss, autograd, AdamW) == the reference modules' own autograd."""
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "train32.npz"))
    loss, grads, new, _ = torch_ref.dose_pyfer_train_step(dose_sd32, vol["dose_input"], vol["gt"], lr=1e-4, weight_decay=1e-4)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    names, norms = list(g["names"]), g["norms"]
    assert sorted(names) == sorted(grads)
    gmax = norms.max()
    for n, want in zip(names, norms):
        got = float(grads[n].double().norm())
        assert abs(got - want) <= 2e-3 * want + 1e-6 * gmax, n
        idx = torch_ref.sample_idx(grads[n].numel())
        ref = torch.from_numpy(g["g/" + n])
        assert (grads[n].flatten()[idx] - ref).abs().max() <= 5e-3 * ref.abs().max() + 1e-6 * gmax, n
    bn = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.1."
    assert torch_ref.rel_l2(new[bn + "running_mean"], torch.from_numpy(g["running_mean"])) < 1e-4
    assert torch_ref.rel_l2(new[bn + "running_var"], torch.from_numpy(g["running_var"])) < 1e-4
    w = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.0.weight"
    idx = torch_ref.sample_idx(new[w].numel())
    assert (new[w].flatten()[idx] - torch.from_numpy(g["p/" + w])).abs().max() < 2.5e-4      # lr * O(1) per Adam step


def test_unfrozen_train_step_matches_reference_fixture(dose_sd32):
    """freeze=False (train_light_pyfer.py:61-88; GenLoss freez=False, loss.py:114-115): the oracle's step == the reference
    modules' own autograd with every parameter trainable (gradients through net_A: strided convs, trilinear upsampling)."""
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "train32_unfrozen.npz"))
    loss, grads, new, _ = torch_ref.dose_pyfer_train_step(dose_sd32, vol["dose_input"], vol["gt"], lr=1e-4, weight_decay=1e-4,
                                                          freeze=False)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    names, norms = list(g["names"]), g["norms"]
    assert any(n.startswith("net_A.") for n in names) and "conv_out_A.weight" in names
    gmax = norms.max()
    for n, want in zip(names, norms):
        got = float(grads[n].double().norm())
        assert abs(got - want) <= 2e-3 * want + 1e-6 * gmax, n
        idx = torch_ref.sample_idx(grads[n].numel())
        ref = torch.from_numpy(g["g/" + n])
        assert (grads[n].flatten()[idx] - ref).abs().max() <= 5e-3 * ref.abs().max() + 1e-6 * gmax, n
    w = "net_A.encoder.encoder_2.0.single_conv.0.weight"          # a stride-2 convolution of net_A
    idx = torch_ref.sample_idx(new[w].numel())
    assert (new[w].flatten()[idx] - torch.from_numpy(g["p/" + w])).abs().max() < 2.5e-4


def test_seg_train_step_matches_reference_fixture(seg_sd32):
    """oracle seg train step (train-mode BN, DiceCE, autograd) == the reference module's own autograd."""
    vol = synth.make_batch(2, 32, seed=1234)
    g = np.load(os.path.join(GOLDEN, "segtrain32.npz"))
    loss, grads, _, _ = torch_ref.oar_transeg_train_step(seg_sd32, vol["ct"], synth.oar_labels(vol["oars"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    names, norms = list(g["names"]), g["norms"]
    assert sorted(names) == sorted(grads)
    gmax = norms.max()
    for n, want in zip(names, norms):
        assert abs(float(grads[n].double().norm()) - want) <= 2e-3 * want + 1e-6 * gmax, n
        ref = torch.from_numpy(g["g/" + n])
        got = grads[n].flatten()[torch_ref.sample_idx(grads[n].numel())]
        assert (got - ref).abs().max() <= 5e-3 * ref.abs().max() + 1e-6 * gmax, n


def _eval_inputs():
    v = synth.make_volume(64, seed=99)
    gt_gy = (v["gt"][0, 0].numpy() * 70).astype(np.float32)
    pmask = v["gt"][0, 1].numpy()
    raw = (gt_gy / 70 + np.random.default_rng(7).normal(0, 0.03, gt_gy.shape)).astype(np.float32)
    return v, raw, gt_gy, pmask


def test_evaluation_matches_reference_fixture():
    """oracle/eval_ref.py == the reference's own evaluate_openKBP.py functions (fixture made from them)."""
    from oracle import eval_ref
    v, raw, gt_gy, pmask = _eval_inputs()
    g = np.load(os.path.join(GOLDEN, "eval64.npz"))
    pred = eval_ref.postprocess(raw, pmask)
    st = {k: m[0, 0].numpy() for k, m in synth.structures(v).items()}
    out = eval_ref.evaluate(pred, gt_gy, pmask, st, (3.906, 3.906, 2.5))
    assert abs(out["dose_dif"] - float(g["dose_dif"])) < 1e-6
    assert np.allclose(out["ivs"], g["ivs"], rtol=0, atol=1e-12)
    assert abs(out["dvh_dif"] - float(g["dvh_dif"])) < 1e-5
    keys = [k for k in g.files if k.startswith("pre") or k.startswith("gt_")]
    assert sorted(keys) == sorted(out["table"]) and len(keys) == 2 * (7 * 2 + 3 * 4)
    for k in keys:
        assert abs(out["table"][k] - float(g[k])) <= 1e-5 * max(1.0, abs(float(g[k]))), k


def test_input_pipeline_matches_reference_fixture():
    """oracle/pipeline_ref.py == the reference's own transform classes (fixture made by running them)."""
    from oracle import pipeline_ref
    rng = np.random.default_rng(11)
    shp = (12, 10, 16)
    raw = {"CT": rng.integers(-2000, 3000, shp).astype(np.int16), "dose": (rng.random(shp) * 75).astype(np.float32),
           "dose_mask": (rng.random(shp) < 0.6).astype(np.uint8)}
    for n in ("Brainstem", "SpinalCord", "LeftParotid", "Mandible", "PTV70", "PTV56"):
        raw[n] = (rng.random(shp) < 0.2).astype(np.uint8)
    g = np.load(os.path.join(GOLDEN, "pipeline12.npz"))
    inp, gt = pipeline_ref.prepare(raw)
    assert np.array_equal(inp, g["input"]) and np.array_equal(gt, g["gt"])


def test_sliding_window_matches_fixture(seg_sd32):
    ct48 = synth.make_volume(48, seed=77)["ct"]
    g = torch.from_numpy(np.load(os.path.join(GOLDEN, "sliding48.npz"))["logits"])
    with torch.no_grad():
        out = torch_ref.sliding_window_logits(seg_sd32, ct48, roi=32, sw_batch=4)
    assert torch_ref.rel_l2(out[:, :, ::2, ::2, ::2], g) < TOL


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_restatement_equals_live_reference_modules():
    torch.manual_seed(3)
    m = ref_loader.build_dose(32, act="relu").eval()
    sd = synth_ckpt.make_state_dict(synth_ckpt.manifest_of(m), seed=5)
    m.load_state_dict(sd, strict=True)
    x = synth.make_batch(1, 32, seed=9)["dose_input"]
    with torch.no_grad():
        want = m(x)
        got = torch_ref.dose_pyfer_forward(sd, x, act="relu")
    assert torch_ref.rel_l2(got[1][0], want[1][0]) < 1e-6
    m2 = ref_loader.build_dose(32, multiS_conv=False).eval()
    sd2 = synth_ckpt.make_state_dict(synth_ckpt.manifest_of(m2), seed=6)
    m2.load_state_dict(sd2, strict=True)
    with torch.no_grad():
        assert torch_ref.rel_l2(torch_ref.dose_pyfer_forward(sd2, x, multiS_conv=False)[1][0], m2(x)[1][0]) < 1e-6


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_manifests_are_current():
    got = [tuple(e) for e in synth_ckpt.manifest_of(ref_loader.build_seg(128))]
    assert got == load_manifest("oar_transeg")


def test_transeg_old_matches_reference_fixture():
    man = [(k, ([1, 8, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("transeg_old_96")]
    sd = synth_ckpt.make_state_dict(man, seed=2)
    vol = synth.make_batch(2, 32, seed=1234)
    g = torch.from_numpy(np.load(os.path.join(GOLDEN, "seg_old32.npz"))["logits"])
    with torch.no_grad():
        out = torch_ref.oar_transeg_forward(sd, vol["ct"][:1], old=True)
    assert torch_ref.rel_l2(out, g) < TOL
