"""Adam8bit-compatible optimizer state: code book, block-wise (de)quantisation, export / import round trip."""
import numpy as np
import pytest
import torch


def test_dynamic_map_has_the_bnb_code_book_shape():
    from dose_prediction_b200 import optim8bit
    for signed in (True, False):
        q = optim8bit.create_dynamic_map(signed)
        assert q.shape == (256,) and bool((q[1:] > q[:-1]).all())          # 256 distinct values, sorted
        assert float(q.max()) == 1.0 and 0.0 in q.tolist()
        assert float(q.min()) == (0.0 if not signed else float(q.min())) and (signed or float(q.min()) == 0.0)
    qs, qu = optim8bit.create_dynamic_map(True), optim8bit.create_dynamic_map(False)
    # signed: 127 magnitudes mirrored, plus 0 and +1; unsigned: 254 magnitudes (twice the resolution), plus 0 and 1
    pos, neg = sorted(v for v in qs.tolist() if v > 0), sorted(-v for v in qs.tolist() if v < 0)
    assert len(neg) == 127 and pos[:-1] == pytest.approx(neg, rel=1e-6) and pos[-1] == 1.0
    assert int((qu > 0).sum()) == 255
    # smallest magnitude 10^-6 * 0.55 (exponent 0: one fraction item, the mean of [0.1, 1]), largest below 1: ~0.9929
    assert pos[0] == pytest.approx(0.55e-6, rel=1e-5) and 0.99 < pos[-2] < 1.0


def test_numpy_restatement_round_trip_error_is_bounded_by_the_code_spacing():
    from dose_prediction_b200 import optim8bit
    from oracle import bnb_ref
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(3 * 2048 + 100) * 1e-3).astype(np.float32)
    q = optim8bit.create_dynamic_map(True).numpy()
    codes, absmax = bnb_ref.quantize_blockwise(x, q)
    assert codes.shape == x.shape and absmax.shape == (4,)
    y = bnb_ref.dequantize_blockwise(codes, absmax, q)
    big = np.abs(x) > 0.1 * np.repeat(absmax, 2048)[:x.size]
    assert np.abs(y - x)[big].max() <= 0.04 * np.repeat(absmax, 2048)[:x.size][big].max()
    assert np.all(np.sign(y[big]) == np.sign(x[big]))


@pytest.mark.gpu
def test_cuda_quantisation_matches_the_numpy_restatement():
    from dose_prediction_b200 import optim8bit
    from oracle import bnb_ref
    torch.manual_seed(3)
    for signed, n in ((True, 5 * 2048 + 77), (False, 4096), (True, 2048 * 300 + 1)):
        x = torch.randn(n, device="cuda:0") * 1e-2
        if not signed:
            x = x * x
        q = optim8bit.create_dynamic_map(signed)
        codes, absmax = optim8bit.quantize_blockwise(x, q)
        rc, ra = bnb_ref.quantize_blockwise(x.cpu().numpy(), q.numpy())
        assert np.array_equal(absmax.cpu().numpy(), ra)
        same = (codes.cpu().numpy() == rc)
        assert same.mean() > 0.9999          # ties between two codes may round differently (fp32 1/absmax vs division)
        y = optim8bit.dequantize_blockwise(codes, absmax, q)
        assert np.allclose(y.cpu().numpy(), bnb_ref.dequantize_blockwise(codes.cpu().numpy(), ra, q.numpy()), rtol=0, atol=0)


@pytest.mark.gpu
def test_export_import_resumes_training_within_the_8bit_rounding():
    """train 3 steps -> export in bnb's Adam8bit layout -> fresh trainer imports it -> the next step's update agrees with the
    uninterrupted run to the 8-bit state's precision; the exported dict has bnb's keys / dtypes / shapes."""
    from conftest import load_manifest
    from dose_prediction_b200 import networks, optim8bit, synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import synth_ckpt
    S, B = 32, 2
    man = [(k, ([1, 8, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
    sd = synth_ckpt.make_state_dict(man, seed=0)
    vol = synth.make_batch(B, S, seed=1234)
    x, gt = vol["dose_input"].cuda(), vol["gt"].cuda()

    def fresh():
        m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(S,) * 3)
        m.load_state_dict(sd, strict=True)
        return m.cuda().train()
    m1 = fresh()
    t1 = DoseTrainer(m1, B, S, lr=1e-3, weight_decay=1e-4)
    for _ in range(3):
        t1.step(x, gt)
    exported = optim8bit.export_state(t1)
    params = list(m1.parameters())
    big = next(i for i, p in enumerate(params) if p.numel() >= 4096 and i in exported["state"])
    st = exported["state"][big]
    assert st["state1"].dtype == torch.uint8 and st["state1"].shape == params[big].shape and st["step"] == 3
    assert st["qmap1"].shape == (256,) and st["absmax1"].shape == ((params[big].numel() + 2047) // 2048,)
    small = next(i for i, p in enumerate(params) if p.numel() < 4096 and i in exported["state"])
    assert exported["state"][small]["state1"].dtype == torch.float32
    assert exported["param_groups"][0]["params"] == list(range(len(params)))
    weights3 = {k: v.clone() for k, v in m1.state_dict().items()}
    t1.step(x, gt)                                                     # uninterrupted 4th step
    # resume: same weights, state from the 8-bit export
    m2 = fresh()
    m2.load_state_dict(weights3, strict=True)
    t2 = DoseTrainer(m2, B, S, lr=1e-3, weight_decay=1e-4)
    assert optim8bit.import_state(t2, exported) == 3
    t2.step(x, gt)
    torch.cuda.synchronize()
    w = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.0.weight"
    d1 = (m1.state_dict()[w] - weights3[w]).flatten()
    d2 = (m2.state_dict()[w] - weights3[w]).flatten()
    assert float((d1 - d2).norm() / d1.norm()) < 0.1                   # 8-bit moments: a few % on the update, not on the weight
    assert float((m1.state_dict()[w] - m2.state_dict()[w]).abs().max()) < 1e-3    # never more than one Adam step (lr)
