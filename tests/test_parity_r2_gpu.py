"""Round-2 parity tests (VERDICT r1 "pin parity on what is benchmarked, and widen the margin").

  * the BENCHMARKED configuration itself — the batch-8 128^3 cascade plan bench.py times — against the oracle;
  * the 128^3 seg net over several weight seeds / volumes, every one >= 99.9 % argmax agreement;
  * BASELINE.json configs[4] (192^3, 1728 ViT tokens, ragged key blocks) and configs[3] (training step at 128^3, batch 2);
  * gradients from a linear probe loss against the fp32 oracle and against the oracle run under the CUDA path's
    operand-rounding recipe (oracle EMU); the test's docstring says what bounds them (flipped ReLU masks);
  * eval-forward -> trainer.step() -> eval-forward (stale inference plans), two devices in one process.

All through the public nn.Module / trainer API, i.e. through the C ABI.  The oracle is oracle/torch_ref.py (fp32); for
the two 128^3-and-larger training / 192^3 cases it is evaluated with the same fp32 arithmetic on the GPU (TF32 off),
which is the "plain PyTorch fp32 reference" of the same op — the CPU run of those sizes needs minutes and > 100 GB.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import load_manifest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sd(name, size, seed):
    from oracle import synth_ckpt
    tokens = (size // 16) ** 3
    man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest(name)]
    return synth_ckpt.make_state_dict(man, seed=seed)


def _dose_model(size, sd, dev=DEV):
    from dose_prediction_b200 import networks
    m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3).eval()
    m.load_state_dict(sd, strict=True)
    return m.to(dev)


def _seg_model(size, sd, dev=DEV):
    from dose_prediction_b200 import networks
    m = networks.OARTranseg(1, 8, (size,) * 3, pos_embed="perceptron").eval()
    m.load_state_dict(sd, strict=True)
    return m.to(dev)


def _rel(a, b):
    from oracle import torch_ref
    return torch_ref.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())


def _fp32_gpu():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def _to(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


@pytest.mark.parametrize("wseed,vseed", [(20, 1234), (31, 1241), (42, 1248)])
def test_seg_128_argmax_margin_on_every_seed(wseed, vseed):
    """north_star: >= 99.9 % voxel-identical argmax and logits rel-L2 <= 1e-2 — on every weight draw / volume."""
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    ssd = _sd("oar_transeg", 128, wseed)
    vol = synth.make_volume(128, seed=vseed)
    with torch.no_grad():
        want = torch_ref.oar_transeg_forward(ssd, vol["ct"])
    got = _seg_model(128, ssd)(vol["ct"].to(DEV)).cpu()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    print("SEED", wseed, vseed, "logits", _rel(got, want), "argmax", agree)
    assert _rel(got, want) < 1e-2
    assert agree >= 0.999


def test_benchmarked_batch8_128_cascade_matches_oracle():
    """The plan bench.py times (batch 8 x 128^3, seg -> hand-off -> dose), entries 0 and 5 against the oracle."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.cascade import CascadePlan
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    B, S = 8, 128
    ssd, dsd = _sd("oar_transeg", S, 20), _sd("dose_pyfer", S, 10)
    vols = synth.make_batch(B, S, seed=1234)
    casc = CascadePlan(_seg_model(S, ssd), _dose_model(S, dsd), B, S, DEV, keep_structures=True, graph=True)
    dose = casc(vols["ct"].to(DEV), vols["ptv"].to(DEV))
    torch.cuda.synchronize()
    casc.plan.check_device_errors()
    for i in (0, 5):
        with torch.no_grad():
            logits = torch_ref.oar_transeg_forward(ssd, vols["ct"][i:i + 1])
            st = torch_ref.handoff(logits, vols["ptv"][i:i + 1], vols["ct"][i:i + 1])
            got_st = casc.structures[i:i + 1].cpu()
            want = torch_ref.dose_pyfer_forward(dsd, got_st)[1][0]
        got_logits = casc.logits[i:i + 1].cpu()
        agree = (got_logits.argmax(1) == logits.argmax(1)).float().mean().item()
        rep = {"entry": i, "logits": _rel(got_logits, logits), "argmax": agree,
               "structures": (got_st == st).float().mean().item(), "dose": _rel(dose[i:i + 1], want)}
        print("BATCH8", rep)
        assert rep["logits"] < 1e-2 and rep["argmax"] >= 0.999 and rep["dose"] < 1e-2


def test_networks_192_match_oracle():
    """BASELINE.json configs[4]: 192^3 (12^3 = 1728 tokens: ragged last key block in the fused attention, 3 x 3 x 3 patch
    grid per 48-voxel... every level 192 / 96 / 48 / 24 / 12)."""
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    _fp32_gpu()
    S = 192
    ssd, dsd = _sd("oar_transeg", S, 20), _sd("dose_pyfer", S, 10)
    vol = synth.make_volume(S, seed=77)
    with torch.no_grad():
        want_logits = torch_ref.oar_transeg_forward(_to(ssd, DEV), vol["ct"].to(DEV)).cpu()
        want = torch_ref.dose_pyfer_forward(_to(dsd, DEV), vol["dose_input"].to(DEV))
        want = [want[0].cpu(), [w.cpu() for w in want[1]]]
    torch.cuda.empty_cache()
    seg = _seg_model(S, ssd)
    logits = seg(vol["ct"].to(DEV)).cpu()
    del seg
    torch.cuda.empty_cache()
    out = _dose_model(S, dsd)(vol["dose_input"].to(DEV))
    torch.cuda.synchronize()
    agree = (logits.argmax(1) == want_logits.argmax(1)).float().mean().item()
    print("P192", _rel(logits, want_logits), agree, _rel(out[0], want[0]), [_rel(a, b) for a, b in zip(out[1], want[1])])
    assert _rel(logits, want_logits) < 1e-2 and agree >= 0.999
    assert _rel(out[0], want[0]) < 1e-2 and all(_rel(a, b) < 1e-2 for a, b in zip(out[1], want[1]))


def _train_model(size, seed=0):
    sd = _sd("dose_pyfer", size, seed)
    from dose_prediction_b200 import networks
    m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).train(), sd


def test_training_step_128_batch2_matches_oracle():
    """BASELINE.json configs[3] at its real size: one Pyfer.training_step + AdamW, 128^3, batch 2."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    _fp32_gpu()
    S, B = 128, 2
    model, sd = _train_model(S)
    vol = synth.make_batch(B, S, seed=1234)
    loss_ref, grads_ref, new_ref, outs_ref = torch_ref.dose_pyfer_train_step(_to(sd, DEV), vol["dose_input"].to(DEV),
                                                                             vol["gt"].to(DEV), lr=1e-4, weight_decay=1e-4)
    grads_ref = {k: v.cpu() for k, v in grads_ref.items()}
    new_ref = {k: v.cpu() for k, v in new_ref.items()}
    outs_ref = [outs_ref[0].cpu(), [o.cpu() for o in outs_ref[1]]]
    loss_ref = loss_ref.cpu()
    torch.cuda.empty_cache()
    tr = DoseTrainer(model, B, S, lr=1e-4, weight_decay=1e-4)
    loss = tr.step(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.check_health()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    for a, b in zip(tr.outputs()[1], outs_ref[1]):
        assert _rel(a, b) < 1e-2
    g = tr.grads()
    worst = (1.0, "")
    for n in ("net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.0.weight",
              "net_B.decoder.decoder1.conv_block.cov_.conv_3.0.conv.3.weight",
              "net_B.decoder.decoder2.conv_block.cov_.conv_7.0.conv.3.weight",
              "net_B.decoder.decoder4.conv_block.cov_.conv_7.0.conv.0.weight",
              "net_B.encoder.skip1.layer.conv1.conv.weight",
              "net_B.encoder.vit.blocks.7.mlp.linear1.weight",
              "net_B.encoder.vit.patch_embedding.patch_embeddings.1.weight",
              "net_B.decoder.decoder3.transp_conv.conv.weight",
              "net_B.dose_convertors.0.0.weight"):
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), grads_ref[n].flatten().double(), dim=0))
        worst = min(worst, (cos, n))
        assert cos > 0.99, (n, cos)
        assert abs(float(g[n].norm()) / float(grads_ref[n].norm()) - 1.0) < 0.06, n
    print("TRAIN128 loss", float(loss), float(loss_ref), "worst cosine", worst)
    bn = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.1."
    after = model.state_dict()
    assert _rel(after[bn + "running_mean"], new_ref[bn + "running_mean"]) < 1e-3
    assert _rel(after[bn + "running_var"], new_ref[bn + "running_var"]) < 1e-3
    assert int(after[bn + "num_batches_tracked"]) == 1          # the reference's BatchNorm3d counts its train-mode forwards


def test_gradients_against_the_oracle_with_a_linear_probe_loss():
    """Backward-pass parity from a LINEAR probe loss  dL/dpred_i = R_i  (no sign() of an L1 loss), once against the plain
    fp32 oracle and once against the oracle run under the operand-rounding recipe of oracle/precision_probe.py (EMU).

    Measured on B200 (scripts/gpu_grad_diag.py): the heads' gradients (no nonlinearity between them and the loss) agree
    to 5e-4..2e-3 — that is the fp16 rounding of the tensor-core gradient operands.  One ReLU further down the error is
    5e-3, and 2e-2 after the BatchNorm+ReLU pairs of the 7^3 branch, 4e-2 at the patch embedding: the forward passes
    differ by 1e-3 (north_star's tolerance is 1e-2), which flips ~1e-3 of the ReLU / LeakyReLU masks, and a flipped mask
    changes that element's gradient by 100 % => O(sqrt(1e-3)) in relative L2.  Running the oracle under EMU does not
    remove this: the recipe reproduces the MAGNITUDE of the CUDA path's rounding, not its bits (accumulation order,
    train-mode BatchNorm statistics), so the two forwards are as far from each other as from the exact one.  The tight
    bounds on the backward arithmetic itself are in tests/test_training_gpu.py (every backward kernel vs fp64 autograd
    on identically rounded operands, 1e-5 .. 1e-3)."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import precision_probe, torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    S, B = 32, 2
    model, sd = _train_model(S)
    vol = synth.make_batch(B, S, seed=1234)
    gen = torch.Generator().manual_seed(5)
    probe = [torch.randn(B, 1, S >> i, S >> i, S >> i, generator=gen) / (S >> i) ** 1.5 for i in range(4)]
    refs = {}
    for tag, emu in (("fp32", None), ("emu", precision_probe.recipe_dose)):
        torch_ref.EMU = emu
        try:
            refs[tag] = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], probe=probe)[1]
        finally:
            torch_ref.EMU = None
    tr = DoseTrainer(model, B, S, probe=probe)
    tr.forward_backward(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.check_health()
    g = tr.grads()
    for tag, grads_ref in refs.items():
        gmax = max(float(v.norm()) for v in grads_ref.values())
        rels = sorted(((_rel(g[n], ref), n) for n, ref in grads_ref.items() if float(ref.norm()) >= 1e-4 * gmax), reverse=True)
        print("GRAD", tag, "worst", rels[:3], "median", rels[len(rels) // 2])
        assert len(rels) > 120
        assert rels[0][0] < 8e-2, rels[:5]                     # every tensor
        assert rels[len(rels) // 2][0] < 4e-2, rels[len(rels) // 2]
        for i in range(4):                                      # mask-free gradients: fp16 operand rounding only
            assert _rel(g[f"net_B.dose_convertors.{i}.0.weight"], grads_ref[f"net_B.dose_convertors.{i}.0.weight"]) < 5e-3
            assert _rel(g[f"net_B.dose_convertors.{i}.0.bias"], grads_ref[f"net_B.dose_convertors.{i}.0.bias"]) < 1e-5


def test_eval_forward_after_trainer_step_sees_the_new_weights():
    """train -> model.eval(); model(x) -> train -> model(x) (the reference's validate-every-N-epochs flow): the trainer
    updates parameters and BatchNorm running statistics through raw pointers, so the cached inference plan must be
    invalidated — the second validation must match the oracle for the UPDATED weights."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    S, B = 32, 2
    model, sd = _train_model(S)
    vol = synth.make_batch(B, S, seed=1234)
    x = vol["dose_input"].to(DEV)
    model.eval()
    y0 = model(x)[1][0].clone()
    model.train()
    tr = DoseTrainer(model, B, S, lr=1e-2, weight_decay=1e-4)          # large lr: the step must be visible
    for _ in range(3):
        tr.step(x, vol["gt"].to(DEV))
    torch.cuda.synchronize()
    model.eval()
    y1 = model(x)[1][0].clone()
    with torch.no_grad():
        want = torch_ref.dose_pyfer_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, vol["dose_input"])[1][0]
    assert _rel(y1, y0) > 1e-2, "inference output did not move after three optimizer steps (stale plan?)"
    assert _rel(y1, want) < 1e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_driven_from_one_process():
    """kernel function attributes (max dynamic shared memory) are per device: a second GPU in the same process must
    work (a process-wide `static bool configured` made it launch with the 48 KB default)."""
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    ssd = _sd("oar_transeg", 32, 1)
    vol = synth.make_volume(32, seed=3)
    with torch.no_grad():
        want = torch_ref.oar_transeg_forward(ssd, vol["ct"])
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            outs.append(_seg_model(32, ssd, dev)(vol["ct"].to(dev)).cpu())
    assert _rel(outs[0], want) < 1e-2 and _rel(outs[1], want) < 1e-2
    assert torch.equal(outs[0], outs[1])


def test_train_mode_forward_is_an_autograd_node_like_the_reference():
    """Pyfer.training_step unchanged (train_light_pyfer.py:122-143,194-197): `output = model(input_)` in train mode, a torch
    loss on the outputs, `loss.backward()`, a torch optimizer step — against the oracle's autograd + AdamW."""
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    S, B = 32, 2
    model, sd = _train_model(S)
    for n, p in model.named_parameters():                       # Pyfer(freeze=True), train_light_pyfer.py:85-88
        if "net_A" in n or "conv_out_A" in n:
            p.requires_grad = False
    vol = synth.make_batch(B, S, seed=1234)
    loss_ref, grads_ref, new_ref, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], lr=1e-4, weight_decay=1e-4)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    out = model(vol["dose_input"].to(DEV))
    assert out[1][0].requires_grad and not out[0].requires_grad
    loss = torch_ref.gen_loss(out, vol["gt"].to(DEV), 10.0, 8.0)          # plain torch ops on the outputs (loss.py:69-119)
    opt.zero_grad()
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    gmax = max(float(v.norm()) for v in grads_ref.values())
    named = dict(model.named_parameters())
    checked = 0
    for n, ref in grads_ref.items():
        g = named[n].grad
        if float(ref.norm()) < 1e-4 * gmax:
            assert g is None or float(g.norm()) < 1e-3 * gmax, n          # unused parameters keep grad None, like the reference
            continue
        cos = float(F.cosine_similarity(g.flatten().double().cpu(), ref.flatten().double(), dim=0))
        assert cos > 0.995, (n, cos)
        checked += 1
    assert checked > 120
    assert named["net_B.out.0.weight"].grad is None and named["net_A.encoder.encoder_1.0.single_conv.0.weight"].grad is None
    opt.step()
    after = model.state_dict()
    w = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.0.weight"
    mine, want = (after[w] - before[w]).flatten().cpu(), (new_ref[w] - sd[w]).flatten()
    assert float((mine.sign() == want.sign()).float().mean()) > 0.97 and float(mine.abs().max()) < 1.2e-4
    assert torch.equal(after["net_B.out.0.weight"], before["net_B.out.0.weight"])           # never touched, never decayed
    # a second forward/backward through the cached training plan; then eval() sees the updated weights
    loss2 = torch_ref.gen_loss(model(vol["dose_input"].to(DEV)), vol["gt"].to(DEV), 10.0, 8.0)
    loss2.backward()
    sd2 = dict(sd)
    sd2.update(new_ref)
    loss2_ref = torch_ref.dose_pyfer_train_step(sd2, vol["dose_input"], vol["gt"], lr=1e-4, weight_decay=1e-4)[0]
    # (the first Adam step RAISES the loss of this random-init net, 8.54 -> 9.55, in the oracle too: what must agree is the value)
    assert abs(float(loss2) - float(loss2_ref)) <= 5e-3 * abs(float(loss2_ref)), (float(loss2), float(loss2_ref))
    model.eval()
    with torch.no_grad():
        y = model(vol["dose_input"].to(DEV))[1][0]
        want_y = torch_ref.dose_pyfer_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, vol["dose_input"])[1][0]
    assert _rel(y, want_y) < 1e-2


def test_seg_train_mode_forward_is_an_autograd_node():
    from dose_prediction_b200 import networks, synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    S, B = 32, 2
    sd = _sd("oar_transeg", S, 1)
    model = networks.OARTranseg(1, 8, (S,) * 3, pos_embed="perceptron")
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train()
    vol = synth.make_batch(B, S, seed=1234)
    label = synth.oar_labels(vol["oars"])
    loss_ref, grads_ref, _, logits_ref = torch_ref.oar_transeg_train_step(sd, vol["ct"], label)
    logits = model(vol["ct"].to(DEV))
    assert logits.requires_grad and _rel(logits, logits_ref) < 1e-2
    loss = torch_ref.dice_ce_loss(logits, label.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    named = dict(model.named_parameters())
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked = 0
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:
            continue
        cos = float(F.cosine_similarity(named[n].grad.flatten().double().cpu(), ref.flatten().double(), dim=0))
        assert cos > 0.99, (n, cos)
        checked += 1
    assert checked > 100


def test_sub_block_forwards_match_the_oracle():
    """SURVEY 8(b) lists their forward signatures verbatim: ModifiedUnetrUpBlock.forward(inp, skip) (base_blocks.py:136-141),
    ViTEncoder.forward (dose_pyfer.py:124-144), PyMSCDecoder.forward (:232-239); composed they must reproduce
    MainSubsetModel.forward."""
    from dose_prediction_b200 import synth
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    S = 32
    dsd = _sd("dose_pyfer", S, 0)
    dose = _dose_model(S, dsd)
    vol = synth.make_batch(1, S, seed=1234)
    with torch.no_grad():
        full = dose(vol["dose_input"].to(DEV))
        x25 = torch.cat((torch_ref.c3d_base_unet(dsd, "net_A.", vol["dose_input"]), vol["dose_input"]), 1).to(DEV)
        enc = dose.net_B.encoder(x25)
        assert [tuple(e.shape[1:]) for e in enc] == [(16, S, S, S), (32, S // 2,) * 1 + (S // 2, S // 2), (64, S // 4, S // 4, S // 4),
                                                     (128, S // 8, S // 8, S // 8), (768, S // 16, S // 16, S // 16)]
        dec = dose.net_B.decoder(enc)
        heads = [F.conv3d(d, dose.net_B.dose_convertors[i][0].weight, dose.net_B.dose_convertors[i][0].bias) for i, d in enumerate(dec)]
    for a, b in zip(heads, full[1]):
        assert _rel(a, b) < 5e-3                                   # same kernels; NCDHW fp32 round trips at the block boundaries
    # one up block on its own vs the oracle's functional restatement of base_blocks.py:136-141
    blk = dose.net_B.decoder.decoder1
    inp, skip = torch.randn(1, 32, 8, 8, 8, device=DEV), torch.randn(1, 16, 16, 16, 16, device=DEV)
    with torch.no_grad():
        got = blk(inp, skip)
        pre = "net_B.decoder.decoder1."
        want = torch_ref.modified_unetr_up_block(dsd, pre, inp.cpu(), skip.cpu(), "mish")
    assert tuple(got.shape) == (1, 16, 16, 16, 16)
    assert _rel(got, want) < 1e-2


def test_captured_training_step_replays_the_eager_schedule():
    """DoseTrainer.capture(): the step as CUDA graphs must reproduce the eager launch list bit for bit."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    S, B = 32, 2
    vol = synth.make_batch(B, S, seed=1234)
    x, gt = vol["dose_input"].to(DEV), vol["gt"].to(DEV)
    losses = {}
    for mode in ("eager", "graph"):
        model, _ = _train_model(S)
        tr = DoseTrainer(model, B, S, lr=1e-3, weight_decay=1e-4)
        out = [float(tr.step(x, gt))]
        if mode == "graph":
            tr.capture()
        out += [float(tr.step(x, gt)) for _ in range(4)]
        torch.cuda.synchronize()
        tr.check_health()
        losses[mode] = out
        final = {k: v.clone() for k, v in model.state_dict().items()} if mode == "eager" else final
        if mode == "graph":
            for k, v in model.state_dict().items():
                assert torch.equal(v, final[k]), k
    assert losses["eager"] == losses["graph"], losses


def test_liveness_arena_is_bit_identical_and_smaller(monkeypatch):
    """Plan.compact(): activation / pre-norm buffers overlaid by liveness in one arena — the cascade's outputs must not change
    by a bit (eager and as a CUDA graph), over several replays with different inputs, and the activation memory must shrink"""
    from dose_prediction_b200 import engine, synth
    from dose_prediction_b200.cascade import CascadePlan
    from test_networks_gpu import _dose_model, _sd, _seg_model
    outs, mem = {}, {}
    batches = [synth.make_batch(2, 32, seed=300 + i) for i in range(3)]
    for compact in (False, True):
        monkeypatch.setattr(engine, "COMPACT", compact)
        for graph in (False, True):
            casc = CascadePlan(_seg_model(32, _sd("oar_transeg", 32, 1)), _dose_model(32, _sd("dose_pyfer", 32, 0)), 2, 32, "cuda:0",
                               graph=graph)
            res = []
            for b in batches:
                res.append((casc(b["ct"].cuda(), b["ptv"].cuda()).clone(), casc.logits.clone()))
            torch.cuda.synchronize()
            casc.plan.check_device_errors()
            outs[(compact, graph)] = res
            mem[(compact, graph)] = (casc.plan.bytes_alloc, getattr(casc.plan, "arena_bytes", None))
    for graph in (False, True):
        for (d0, l0), (d1, l1) in zip(outs[(False, graph)], outs[(True, graph)]):
            assert torch.equal(d0, d1) and torch.equal(l0, l1)
    before, after = mem[(True, False)][1]
    print("activation bytes %d -> arena %d" % (before, after))
    assert after < 0.7 * before
    assert mem[(True, False)][0] < mem[(False, False)][0]
