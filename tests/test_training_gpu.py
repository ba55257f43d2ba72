"""Training-step parity on a real B200, every call through the C ABI.

Per-kernel tests compare each backward / loss / optimizer kernel with torch autograd (fp64 on CPU) on the operands
exactly as the kernel sees them.  The end-to-end test compares one DoseTrainer step with the oracle
(oracle.torch_ref.dose_pyfer_train_step, pinned against the reference's own autograd by tests/test_oracle.py).
End-to-end gradient tolerances are looser than the forward's: the network is piecewise linear (ReLU / LeakyReLU
masks, the sign() of the L1 loss), so the ~1e-3 forward difference of the fp16 path flips a ~1e-3 fraction of the
masks and moves gradient tensors by O(sqrt(1e-3)) in relative L2 — the per-kernel tests carry the tight bounds.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_manifest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tplan():
    from dose_prediction_b200.training import TrainPlan
    return TrainPlan(torch.device(DEV), 1.0)


def _c8_to_ncdhw(t, C):
    n, cb, d, h, w, _ = t.shape
    return t.permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, d, h, w)[:, :C]


def _ncdhw_to_c8(x, dtype=torch.float32):
    n, c, d, h, w = x.shape
    cb = (c + 15) // 16 * 2
    buf = torch.zeros(n, cb * 8, d, h, w, device=x.device, dtype=dtype)
    buf[:, :c] = x.to(dtype)
    return buf.view(n, cb, 8, d, h, w).permute(0, 1, 3, 4, 5, 2).contiguous()


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))


def _act_of(P, x):
    """fp32 NCDHW -> Act holding fp16(x)."""
    from dose_prediction_b200.engine import Act
    buf = _ncdhw_to_c8(x, torch.float16)
    P.keep.append(buf)
    return Act(buf, 0, x.shape[1], None)


def _raw_of(P, x, stats=True):
    from dose_prediction_b200.engine import Raw
    t = _ncdhw_to_c8(x)
    P.keep.append(t)
    st = None
    if stats:
        xd = x.double()
        st = torch.stack((xd.sum((2, 3, 4)), (xd * xd).sum((2, 3, 4))), dim=-1).contiguous().view(-1)
        P.keep.append(st)
    return Raw(t, x.shape[1], st)


def _finish(P):
    torch.cuda.synchronize()
    P.check_device_errors()


_ACT = {"relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.01), "mish": F.mish, None: lambda t: t}


@pytest.mark.parametrize("act,res,act2,bn", [
    ("relu", None, None, False),
    ("mish", None, None, False),
    (None, "raw", "lrelu", False),        # UnetResBlock with the 1x1x1 + norm3 residual
    (None, "act", "lrelu", False),        # identity residual
    ("relu", None, None, True),           # train-mode BatchNorm3d
])
def test_norm_act_backward_matches_autograd(act, res, act2, bn):
    torch.manual_seed(1)
    N, C, dims = 2, 24, (6, 10, 12)
    x = torch.randn(N, C, *dims, device=DEV) * 1.7 + 0.3
    e = torch.randn(N, C, *dims, device=DEV)
    dy = torch.randn(N, C, *dims, device=DEV)
    P = _tplan()
    raw = _raw_of(P, x)
    out = P.new_act(N, C, dims)
    kw = {}
    bnm = None
    if bn:
        bnm = torch.nn.BatchNorm3d(C).to(DEV).train()
        with torch.no_grad():
            bnm.weight.uniform_(0.5, 1.5)
            bnm.bias.uniform_(-0.5, 0.5)
        P.grad_of[id(bnm.weight)] = torch.zeros(C, device=DEV)
        P.grad_of[id(bnm.bias)] = torch.zeros(C, device=DEV)
        kw["bn"] = bnm
    r_in = None
    if res == "raw":
        r_in = _raw_of(P, e)
    elif res == "act":
        r_in = _act_of(P, e)
    P.t_norm(raw, out, act=act, res=r_in, act_after_res=act2, **kw)
    gy = _ncdhw_to_c8(dy)
    P.add_act_grad(out, (gy, gy.shape[1], 0))
    P.tape[-1]()
    P.run()
    _finish(P)
    # torch reference in fp64 on the same operands
    xd = x.double().cpu().requires_grad_(True)
    ed = (e.half().double() if res == "act" else e.double()).cpu().requires_grad_(True)
    if bn:
        ref_bn = torch.nn.BatchNorm3d(C).double().train()
        with torch.no_grad():
            ref_bn.weight.copy_(bnm.weight.double().cpu())
            ref_bn.bias.copy_(bnm.bias.double().cpu())
        t = ref_bn(xd)
    else:
        t = F.instance_norm(xd, eps=1e-5)
    y = _ACT[act](t)
    if res == "raw":
        y = _ACT[act2](y + F.instance_norm(ed, eps=1e-5))
    elif res == "act":
        y = _ACT[act2](y + ed)
    y.backward(dy.double().cpu())
    fwd = _c8_to_ncdhw(out.buf.float(), C)
    assert _rel(fwd, y.detach()) < 1e-3                               # fp16 storage of the forward output
    g16 = P.raw_grad[raw.t.data_ptr()]
    assert _rel(_c8_to_ncdhw(g16.buf.float(), C), xd.grad) < 1e-3    # fp16 storage of dx
    if res == "raw":
        r16 = P.raw_grad[r_in.t.data_ptr()]
        assert _rel(_c8_to_ncdhw(r16.buf.float(), C), ed.grad) < 1e-3
    elif res == "act":
        (gt, _, _), = P.act_grads[P._key(r_in)]
        assert _rel(_c8_to_ncdhw(gt, C), ed.grad) < 1e-5
    if bn:
        assert _rel(P.grad(bnm.weight), ref_bn.weight.grad) < 1e-4
        assert _rel(P.grad(bnm.bias), ref_bn.bias.grad) < 1e-4
        assert _rel(bnm.running_mean, ref_bn.running_mean) < 1e-5
        assert _rel(bnm.running_var, ref_bn.running_var) < 1e-5


def test_instance_norm_of_activation_backward():
    """IN applied to an already-activated tensor (conv_3_1: relu(IN(conv)) -> IN -> act), gradient in fp32."""
    torch.manual_seed(2)
    N, C, dims = 2, 16, (4, 8, 8)
    x = torch.randn(N, C, *dims, device=DEV).abs()
    dy = torch.randn(N, C, *dims, device=DEV)
    P = _tplan()
    a = _act_of(P, x)
    xh = a.buf.float()
    xs = _c8_to_ncdhw(xh, C).double()
    st = torch.stack((xs.sum((2, 3, 4)), (xs * xs).sum((2, 3, 4))), dim=-1).contiguous().view(-1)
    out = P.new_act(N, C, dims)
    P.t_norm(a, out, act="mish", stats=st)
    gy = _ncdhw_to_c8(dy)
    P.add_act_grad(out, (gy, gy.shape[1], 0))
    P.tape[-1]()
    P.run()
    _finish(P)
    xd = xs.cpu().requires_grad_(True)
    F.mish(F.instance_norm(xd, eps=1e-5)).backward(dy.double().cpu())
    (gt, _, _), = P.act_grads[P._key(a)]
    assert _rel(_c8_to_ncdhw(gt, C), xd.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,k,dil,dims,tc", [
    (16, 16, 3, 1, (6, 8, 16), True),
    (32, 16, 7, 1, (5, 9, 20), True),
    (25, 16, 3, 1, (4, 8, 8), True),            # two parts (16 + 9 channels), like the encoder1 input
    (16, 32, 7, 1, (9, 40, 80), True),          # several row blocks, two W tiles, two C_out tiles
    (32, 32, 3, 1, (8, 30, 4), True),           # W < 16 (deep decoder levels of small volumes)
    (64, 32, 3, 2, (6, 6, 8), True),            # dilated: CUDA-core kernel
    (16, 16, 3, 1, (6, 8, 16), False),
    (32, 16, 7, 1, (5, 9, 20), False),
])
def test_conv_wgrad_and_dgrad_match_autograd(cin, cout, k, dil, dims, tc, monkeypatch):
    from dose_prediction_b200 import training
    monkeypatch.setattr(training, "WGRAD_TC", tc)
    torch.manual_seed(3)
    N = 2
    x = torch.randn(N, cin, *dims, device=DEV)
    g = torch.randn(N, cout, *dims, device=DEV)
    conv = torch.nn.Conv3d(cin, cout, k, padding=dil * (k // 2), dilation=dil).to(DEV)
    P = _tplan()
    P.grad_of[id(conv.weight)] = torch.zeros_like(conv.weight)
    if cin == 25:
        parts = P.new_concat(N, [16, 9], dims)
        xa, xb = x[:, :16].contiguous(), x[:, 16:].contiguous()
        P.keep += [xa, xb]
        P.pack_input(xa, parts[0])
        P.pack_input(xb, parts[1])
        need = False
    else:
        parts = [P.new_act(N, cin, dims)]
        P.pack_input(x, parts[0])
        need = True
    raw = P.t_conv(parts, conv, k, dil, need_dgrad=need)
    g16 = _act_of(P, g)
    P.raw_grad[raw.t.data_ptr()] = g16
    P.tape[-1]()
    with torch.no_grad():
        conv.weight.mul_(1.5)               # the parameters changed since the plan was built ...
    assert len(P.refresh_launches) == (2 if need else 1) and not P.refresh
    P.refresh_weights()                     # ... and are re-packed on the device (dp_pack_conv_weight)
    P.run()
    _finish(P)
    xd = x.half().double().cpu().requires_grad_(True)
    wd = conv.weight.detach().half().double().cpu().requires_grad_(True)
    y = F.conv3d(xd, wd, conv.bias.detach().double().cpu(), padding=dil * (k // 2), dilation=dil)
    assert _rel(_c8_to_ncdhw(raw.t, cout), y.detach()) < 2e-5
    y.backward(g.half().double().cpu())
    assert _rel(P.grad(conv.weight), wd.grad) < 2e-5
    if need:
        (gt, _, _), = P.act_grads[P._key(parts[0])]
        assert _rel(_c8_to_ncdhw(gt, cin), xd.grad) < 2e-5


def test_pointwise_and_head_backward():
    torch.manual_seed(4)
    N, dims = 2, (4, 6, 8)
    xa, xb = torch.randn(N, 16, *dims, device=DEV), torch.randn(N, 16, *dims, device=DEV)
    conv = torch.nn.Conv3d(32, 16, 1).to(DEV)
    g = torch.randn(N, 16, *dims, device=DEV)
    P = _tplan()
    for p in conv.parameters():
        P.grad_of[id(p)] = torch.zeros_like(p)
    za, zb = P.new_concat(N, [16, 16], dims)
    P.pack_input(xa, za)
    P.pack_input(xb, zb)
    raw = P.t_pointwise([za, zb], conv)
    P.raw_grad[raw.t.data_ptr()] = _act_of(P, g)
    P.tape[-1]()
    head = torch.nn.Conv3d(16, 1, 1).to(DEV)
    for p in head.parameters():
        P.grad_of[id(p)] = torch.zeros_like(p)
    xh = P.new_act(N, 16, dims)
    P.pack_input(xa, xh)
    y = P.t_head(xh, head)
    gp = torch.randn(N, 1, *dims, device=DEV)
    P.planar_grad[y.data_ptr()] = gp
    P.tape[-1]()
    for acc64, gr, n in P.finalizers:
        P.add("dp_grad_finalize", acc64.data_ptr(), gr.data_ptr(), n, 1.0)
    P.run()
    _finish(P)
    xd = torch.cat((xa, xb), 1).half().double().cpu().requires_grad_(True)
    wd, bd = conv.weight.detach().double().cpu().requires_grad_(True), conv.bias.detach().double().cpu().requires_grad_(True)
    F.conv3d(xd, wd, bd).backward(g.half().double().cpu())
    assert _rel(P.grad(conv.weight), wd.grad) < 1e-5
    assert _rel(P.grad(conv.bias), bd.grad) < 1e-5
    (ga, _, oa), = P.act_grads[P._key(za)]
    assert _rel(_c8_to_ncdhw(ga, 32), xd.grad) < 1e-5
    xd2 = xa.half().double().cpu().requires_grad_(True)
    wh, bh = head.weight.detach().double().cpu().requires_grad_(True), head.bias.detach().double().cpu().requires_grad_(True)
    yr = F.conv3d(xd2, wh, bh)
    assert _rel(y, yr.detach()) < 1e-5
    yr.backward(gp.double().cpu())
    assert _rel(P.grad(head.weight), wh.grad) < 1e-5 and _rel(P.grad(head.bias), bh.grad) < 1e-5
    (gh, _, _), = P.act_grads[P._key(xh)]
    assert _rel(_c8_to_ncdhw(gh, 16), xd2.grad) < 1e-5


@pytest.mark.parametrize("tokens", [False, True])
def test_deconv_backward(tokens):
    from dose_prediction_b200.engine import Tokens
    torch.manual_seed(5)
    N, Ci, Co, grid = 2, 32, 16, (2, 3, 4)
    w = torch.nn.Parameter(torch.randn(Ci, Co, 2, 2, 2, device=DEV) * 0.2)
    x = torch.randn(N, Ci, *grid, device=DEV)
    odims = tuple(2 * g for g in grid)
    g1, g2 = torch.randn(N, Co, *odims, device=DEV), torch.randn(N, Co, *odims, device=DEV)
    P = _tplan()
    P.grad_of[id(w)] = torch.zeros_like(w)
    out = P.new_act(N, Co, odims)
    if tokens:
        tok = x.permute(0, 2, 3, 4, 1).reshape(N, -1, Ci).half().contiguous()
        src = Tokens(tok, grid)
        P.keep.append(tok)
    else:
        src = P.new_act(N, Ci, grid)
        P.pack_input(x, src)
    P.t_deconv(src, w, out)
    for g in (g1, g2):                      # two consumers of the deconv output
        gc = _ncdhw_to_c8(g)
        P.add_act_grad(out, (gc, gc.shape[1], 0))
    P.tape[-1]()
    for acc64, gr, n in P.finalizers:
        P.add("dp_grad_finalize", acc64.data_ptr(), gr.data_ptr(), n, 1.0)
    P.run()
    _finish(P)
    xd = x.half().double().cpu().requires_grad_(True)
    wd = w.detach().double().cpu().requires_grad_(True)
    wf = w.detach().half().double().cpu() if tokens else wd      # the token path feeds fp16 weights to the GEMM
    y = F.conv_transpose3d(xd, wd, stride=2)
    y.backward((g1 + g2).double().cpu())
    assert _rel(_c8_to_ncdhw(out.buf.float(), Co), F.conv_transpose3d(xd.detach(), wf, stride=2)) < 1e-3
    assert _rel(P.grad(w), wd.grad) < 1e-5
    if tokens:
        (dt,) = P.tok_grads[src.t.data_ptr()]
        assert _rel(dt.view(N, *grid, Ci).permute(0, 4, 1, 2, 3), xd.grad) < 1e-5
    else:
        (gt, _, _), = P.act_grads[P._key(src)]
        assert _rel(_c8_to_ncdhw(gt, Ci), xd.grad) < 1e-5


def test_genloss_forward_backward_matches_oracle():
    from dose_prediction_b200 import _lib, synth
    from oracle import torch_ref
    lib = _lib.lib()
    vol = synth.make_batch(2, 32, seed=11)
    gt = vol["gt"]
    torch.manual_seed(6)
    preds = [(F.interpolate(gt[:, :1], size=(32 >> i,) * 3, mode="trilinear") + 0.1 * torch.randn(2, 1, *(32 >> i,) * 3))
             .requires_grad_(True) for i in range(4)]
    loss = torch_ref.gen_loss([None, preds], gt, 10.0, 8.0)
    loss.backward()
    s = torch.cuda.current_stream().cuda_stream
    acc = torch.zeros(8, dtype=torch.float64, device=DEV)
    out = torch.zeros(1, device=DEV)
    gtd = gt.to(DEV).contiguous()
    pd = [p.detach().to(DEV).contiguous() for p in preds]
    for i, p in enumerate(pd):
        _lib.check(lib.dp_masked_l1(p.data_ptr(), gtd.data_ptr(), 2, 32, 32 >> i, acc[2 * i:].data_ptr(), 0, 0.0, None, s))
    _lib.check(lib.dp_genloss_finalize(acc.data_ptr(), 4, 10.0, 8.0, None, 0.0, out.data_ptr(), s))
    assert abs(float(out) - float(loss)) <= 1e-5 * abs(float(loss))
    for i, p in enumerate(pd):
        g = torch.empty_like(p)
        coef = 10.0 if i == 0 else 8.0 / 3
        _lib.check(lib.dp_masked_l1(p.data_ptr(), gtd.data_ptr(), 2, 32, 32 >> i, acc[2 * i:].data_ptr(), 1, coef, g.data_ptr(), s))
        torch.cuda.synchronize()
        assert _rel(g, preds[i].grad) < 1e-5


def test_adamw_matches_torch():
    from dose_prediction_b200 import _lib
    lib = _lib.lib()
    torch.manual_seed(7)
    n = 10007
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    found = torch.zeros(1, dtype=torch.int32, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    for step in range(1, 6):
        g = torch.randn(n) * (0.1 if step % 2 else 3.0)
        ref.grad = g.clone()
        opt.step()
        gs = (g * 1024.0).to(DEV)
        _lib.check(lib.dp_adamw(p.data_ptr(), gs.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step,
                                1.0 / 1024.0, found.data_ptr(), s))
    torch.cuda.synchronize()
    assert (p.cpu() - ref.detach()).abs().max() < 2e-6
    # a non-finite gradient skips the update
    before = p.clone()
    gs = torch.full((n,), float("inf"), device=DEV)
    _lib.check(lib.dp_grad_check(gs.data_ptr(), n, found.data_ptr(), s))
    _lib.check(lib.dp_adamw(p.data_ptr(), gs.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, 6,
                            1.0, found.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(found) == 1 and torch.equal(p, before)


def test_token_backward_kernels():
    from dose_prediction_b200 import _lib
    lib = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(8)
    rows, cols = 37, 768
    x = torch.randn(rows, cols, device=DEV) * 2 + 0.5
    gamma = torch.rand(cols, device=DEV) + 0.5
    dy, add = torch.randn(rows, cols, device=DEV), torch.randn(rows, cols, device=DEV)
    dx = torch.empty_like(x)
    dg, db = torch.zeros(cols, dtype=torch.float64, device=DEV), torch.zeros(cols, dtype=torch.float64, device=DEV)
    _lib.check(lib.dp_layernorm_bwd(x.data_ptr(), gamma.data_ptr(), dy.data_ptr(), add.data_ptr(), rows, cols, dx.data_ptr(),
                                    dg.data_ptr(), db.data_ptr(), s))
    xd = x.double().cpu().requires_grad_(True)
    gd = gamma.double().cpu().requires_grad_(True)
    bd = torch.zeros(cols, dtype=torch.float64, requires_grad=True)
    F.layer_norm(xd, (cols,), gd, bd, 1e-5).backward(dy.double().cpu())
    torch.cuda.synchronize()
    assert _rel(dx, xd.grad + add.double().cpu()) < 1e-5
    assert _rel(dg, gd.grad) < 1e-5 and _rel(db, bd.grad) < 1e-5
    # softmax backward
    R, C, ld = 45, 27, 32
    sc = torch.randn(R, C, device=DEV)
    probs = torch.zeros(R, ld, dtype=torch.float16, device=DEV)
    probs[:, :C] = torch.softmax(sc, -1).half()
    dP = torch.randn(R, C, device=DEV)
    dS = torch.zeros(R, ld, dtype=torch.float16, device=DEV)
    _lib.check(lib.dp_softmax_bwd(probs.data_ptr(), ld, dP.data_ptr(), C, R, C, dS.data_ptr(), ld, s))
    pd = probs[:, :C].double().cpu()
    want = pd * (dP.double().cpu() - (pd * dP.double().cpu()).sum(-1, keepdim=True))
    torch.cuda.synchronize()
    assert _rel(dS[:, :C], want) < 1e-3 and float(dS[:, C:].abs().max()) == 0.0
    # GELU forward / backward
    u = torch.randn(5000, device=DEV) * 2
    dh = torch.randn(5000, device=DEV)
    h16 = torch.empty(5000, dtype=torch.float16, device=DEV)
    du = torch.empty(5000, device=DEV)
    _lib.check(lib.dp_act_fwd(u.data_ptr(), 5000, 4, h16.data_ptr(), s))
    _lib.check(lib.dp_act_bwd(u.data_ptr(), dh.data_ptr(), 5000, 4, du.data_ptr(), None, s))
    ud = u.double().cpu().requires_grad_(True)
    yd = F.gelu(ud)
    yd.backward(dh.double().cpu())
    torch.cuda.synchronize()
    assert _rel(h16, yd.detach()) < 1e-3 and _rel(du, ud.grad) < 1e-5
    # batched transpose-cast and head split / merge round trip
    src = torch.randn(3, 10, 21, device=DEV)
    dst = torch.zeros(3, 21, 16, dtype=torch.float16, device=DEV)
    _lib.check(lib.dp_transpose(src.data_ptr(), 1, 10 * 21, 21, 10, 21, dst.data_ptr(), 21 * 16, 16, 3, 2.0, s))
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :, :10], (2.0 * src).transpose(1, 2).half()) and float(dst[:, :, 10:].abs().max()) == 0.0
    B, T, heads, hd = 2, 8, 6, 16
    rowsrc = torch.randn(B, T, 3 * heads * hd, device=DEV)
    split = torch.empty(B * heads, T, hd, device=DEV)
    _lib.check(lib.dp_heads(rowsrc.data_ptr(), 1, split.data_ptr(), 1, B, T, heads, hd, 3 * heads * hd, heads * hd, 0, 1.0, s))
    want = rowsrc[:, :, heads * hd:2 * heads * hd].view(B, T, heads, hd).permute(0, 2, 1, 3).reshape(B * heads, T, hd)
    torch.cuda.synchronize()
    assert torch.equal(split, want)
    back = torch.zeros_like(rowsrc)
    _lib.check(lib.dp_heads(split.data_ptr(), 1, back.data_ptr(), 1, B, T, heads, hd, 3 * heads * hd, heads * hd, 1, 1.0, s))
    torch.cuda.synchronize()
    assert torch.equal(back[:, :, heads * hd:2 * heads * hd], rowsrc[:, :, heads * hd:2 * heads * hd])


def _dose_model(size):
    from dose_prediction_b200 import networks
    from oracle import synth_ckpt
    tokens = (size // 16) ** 3
    man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
    sd = synth_ckpt.make_state_dict(man, seed=0)
    model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(size,) * 3)
    model.load_state_dict(sd, strict=True)
    return model.to(DEV).train(), sd


def test_training_step_matches_oracle_32():
    """One Pyfer.training_step + AdamW update at 32^3, batch 2, vs the oracle's autograd (fp32 CPU)."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    model, sd = _dose_model(32)
    vol = synth.make_batch(2, 32, seed=1234)
    loss_ref, grads_ref, new_ref, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], lr=1e-4,
                                                                             weight_decay=1e-4)
    tr = DoseTrainer(model, 2, 32, lr=1e-4, weight_decay=1e-4)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    loss = tr.step(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    for a, b in zip(tr.outputs()[1], outs_ref[1]):
        assert _rel(a, b) < 1e-2                                   # north_star forward tolerance
    g = tr.grads()
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked = 0
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:                         # analytically-zero gradients (biases feeding a norm)
            assert float(g[n].norm()) < 1e-3 * gmax, n
            continue
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), ref.flatten().double(), dim=0))
        assert cos > 0.995, (n, cos)
        assert abs(float(g[n].norm()) / float(ref.norm()) - 1.0) < 0.05, n
        checked += 1
    assert checked > 120
    after = model.state_dict()
    bn = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.1."
    assert _rel(after[bn + "running_mean"], new_ref[bn + "running_mean"]) < 1e-3
    assert _rel(after[bn + "running_var"], new_ref[bn + "running_var"]) < 1e-3
    # frozen sub-network untouched, trained parameters moved by ~lr in the oracle's direction
    assert torch.equal(after["net_A.encoder.encoder_1.0.single_conv.0.weight"], before["net_A.encoder.encoder_1.0.single_conv.0.weight"])
    w = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.0.weight"
    mine, want = (after[w] - before[w]).flatten().cpu(), (new_ref[w] - sd[w]).flatten()
    assert float((mine.sign() == want.sign()).float().mean()) > 0.97
    assert float(mine.abs().max()) < 1.2e-4


def test_lerp2x_bwd_is_the_transpose_of_the_trilinear_upsample():
    """three dp_lerp2x_bwd passes == autograd of F.interpolate(scale_factor=2, 'trilinear', align_corners=True) (c3d.py:36)"""
    from dose_prediction_b200 import _lib
    lib = _lib.lib()
    torch.manual_seed(5)
    N, C, (D, H, W) = 2, 16, (3, 5, 8)
    g = torch.randn(N, C, 2 * D, 2 * H, 2 * W, device=DEV)
    x = torch.zeros(N, C, D, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True).backward(g.double())
    g8 = _ncdhw_to_c8(g)
    ncb = C // 8
    t1 = torch.zeros(N, ncb, D, 2 * H, 2 * W, 8, device=DEV)
    t2 = torch.zeros(N, ncb, D, H, 2 * W, 8, device=DEV)
    out = torch.zeros(N, ncb, D, H, W, 8, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dp_lerp2x_bwd(g8.data_ptr(), N * ncb, D, 4 * H * W, t1.data_ptr(), s))
    _lib.check(lib.dp_lerp2x_bwd(t1.data_ptr(), N * ncb * D, H, 2 * W, t2.data_ptr(), s))
    _lib.check(lib.dp_lerp2x_bwd(t2.data_ptr(), N * ncb * D * H, W, 1, out.data_ptr(), s))
    torch.cuda.synchronize()
    assert _rel(_c8_to_ncdhw(out, C), x.grad) < 1e-6


def test_unfrozen_training_step_matches_oracle_32():
    """Pyfer(freeze=False) (train_light_pyfer.py:61-88; GenLoss freez=False, loss.py:114-115): every parameter trains, the
    backward pass runs through net_A (stride-2 convs on the space-to-depth copy, trilinear up-sampling, affine InstanceNorm,
    the patch-embedding / res-block data gradients into net_A's output); vs the oracle pinned to the reference's autograd
    (tests/golden/train32_unfrozen.npz)."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    model, sd = _dose_model(32)
    vol = synth.make_batch(2, 32, seed=1234)
    loss_ref, grads_ref, new_ref, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], lr=1e-4,
                                                                             weight_decay=1e-4, freeze=False)
    tr = DoseTrainer(model, 2, 32, lr=1e-4, weight_decay=1e-4, freeze=False)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    loss = tr.step(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    outs = tr.outputs()
    assert _rel(outs[0], outs_ref[0]) < 1e-2
    for a, b in zip(outs[1], outs_ref[1]):
        assert _rel(a, b) < 1e-2
    g = tr.grads()
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked_a = 0
    worst = (1.0, None)
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:
            assert float(g[n].norm()) < 1e-3 * gmax, n
            continue
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), ref.flatten().double(), dim=0))
        worst = min(worst, (cos, n))
        assert cos > 0.99, (n, cos)
        assert abs(float(g[n].norm()) / float(ref.norm()) - 1.0) < 0.06, n
        checked_a += n.startswith("net_A.") or n.startswith("conv_out_A")
    print("worst gradient cosine", worst)
    assert checked_a >= 2 * 19 + 1            # 19 convs of net_A (weights) + their affine InstanceNorms, conv_out_A
    after = model.state_dict()
    for w in ("net_A.encoder.encoder_2.0.single_conv.0.weight", "net_A.decoder.upconv_1.conv.0.weight", "conv_out_A.weight"):
        mine, want = (after[w] - before[w]).flatten().cpu(), (new_ref[w] - sd[w]).flatten()
        assert float((mine.sign() == want.sign()).float().mean()) > 0.95, w
        assert 0.0 < float(mine.abs().max()) < 1.2e-4, w


def test_unfrozen_autograd_forward_trains_net_a():
    """train-mode model(x) with every parameter requiring grad (freeze=False under the unchanged reference training code):
    out_A is differentiable, .grad lands on net_A's parameters and equals the fused trainer's gradient"""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    model, sd = _dose_model(32)
    vol = synth.make_batch(2, 32, seed=1234)
    x, gt = vol["dose_input"].to(DEV), vol["gt"].to(DEV)
    out = model(x)
    assert out[0].requires_grad
    loss = torch_ref.gen_loss(out, gt, 10.0, 8.0, freeze=False)
    loss.backward()
    ga = model.net_A.encoder.encoder_3[0].single_conv[0].weight.grad.clone()
    gc = model.conv_out_A.weight.grad.clone()
    assert float(ga.norm()) > 0 and float(gc.norm()) > 0
    model2, _ = _dose_model(32)
    tr = DoseTrainer(model2, 2, 32, freeze=False)
    loss2 = tr.forward_backward(x, gt)
    g2 = tr.grads()
    assert abs(float(loss) - float(loss2)) <= 1e-4 * abs(float(loss2))
    assert _rel(ga, g2["net_A.encoder.encoder_3.0.single_conv.0.weight"]) < 2e-3
    assert _rel(gc, g2["conv_out_A.weight"]) < 2e-3


@pytest.mark.parametrize("variant", ["dual_dilated_relu", "unetr_up_block"])
def test_decoder_variants_training_step_matches_oracle_32(variant):
    """the other two decoders the constructor flags reach, in train mode: multiS_conv=False (DualDilatedBlock: dilation
    1 / 2 / 3 branches, blocks_MDUNet.py:194-215) with act='relu', and mode_multi_dec=False (monai UnetrUpBlock)."""
    from dose_prediction_b200 import networks, synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import synth_ckpt, torch_ref
    kw_model, kw_ref = (dict(multiS_conv=False, act="relu"), dict(multiS_conv=False, act="relu")) \
        if variant == "dual_dilated_relu" else (dict(mode_multi_dec=False), {})
    model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(32,) * 3, **kw_model)
    sd = synth_ckpt.make_state_dict(synth_ckpt.manifest_of(model), seed=7)
    model.load_state_dict(sd, strict=True)
    model.to(DEV).train()
    vol = synth.make_batch(2, 32, seed=1234)
    loss_ref, grads_ref, _, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], **kw_ref)
    tr = DoseTrainer(model, 2, 32)
    loss = tr.step(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    for a, b in zip(tr.outputs()[1], outs_ref[1]):
        assert _rel(a, b) < 1e-2
    g = tr.grads()
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked = 0
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:
            assert float(g[n].norm()) < 1e-3 * gmax, n
            continue
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), ref.flatten().double(), dim=0))
        # ReLU everywhere + the sign() of the L1 loss: the fp16 forward's 1e-3 difference flips more masks than with Mish
        # (measured worst tensor: 0.988, an encoder deconv behind ~20 nonlinear layers)
        assert cos > 0.98, (n, cos)
        # (the heads' scalar biases are sums of +-1 / count that nearly cancel: absolute slack relative to the largest tensor)
        assert abs(float(g[n].norm()) - float(ref.norm())) < 0.06 * float(ref.norm()) + 1e-3 * gmax, n
        checked += 1
    assert checked > 100


def test_training_reduces_the_loss():
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    model, _ = _dose_model(32)
    vol = synth.make_batch(2, 32, seed=4321)
    tr = DoseTrainer(model, 2, 32, lr=1e-3, weight_decay=1e-4)
    x, gt = vol["dose_input"].to(DEV), vol["gt"].to(DEV)
    losses = [float(tr.step(x, gt)) for _ in range(10)]
    print("losses", losses)
    tr.P.check_device_errors()
    assert all(l == l for l in losses)
    assert losses[-1] < 0.9 * losses[0], losses


def test_dice_ce_forward_backward_matches_oracle():
    from dose_prediction_b200 import _lib, synth
    from oracle import torch_ref
    lib = _lib.lib()
    torch.manual_seed(9)
    N, C, S = 2, 8, 16
    vol = synth.make_batch(N, S, seed=21)
    label = synth.oar_labels(vol["oars"])
    logits = (torch.randn(N, C, S, S, S) * 2).requires_grad_(True)
    loss = torch_ref.dice_ce_loss(logits, label)
    loss.backward()
    s = torch.cuda.current_stream().cuda_stream
    zc = _ncdhw_to_c8(logits.detach().to(DEV))
    lab = label.to(DEV).contiguous()
    acc = torch.zeros(N * 24 + 2, dtype=torch.float64, device=DEV)
    out = torch.zeros(1, device=DEV)
    g16 = torch.zeros(N, 2, S, S, S, 8, dtype=torch.float16, device=DEV)
    vox = S ** 3
    _lib.check(lib.dp_dice_ce(zc.data_ptr(), zc.shape[1], lab.data_ptr(), N, C, vox, acc.data_ptr(), 0, 0.0, None, 0, s))
    _lib.check(lib.dp_dice_ce_finalize(acc.data_ptr(), N, C, vox, out.data_ptr(), s))
    _lib.check(lib.dp_dice_ce(zc.data_ptr(), zc.shape[1], lab.data_ptr(), N, C, vox, acc.data_ptr(), 1, 4096.0, g16.data_ptr(), 2, s))
    torch.cuda.synchronize()
    assert abs(float(out) - float(loss)) <= 1e-5 * abs(float(loss))
    assert _rel(_c8_to_ncdhw(g16.float(), C) / 4096.0, logits.grad) < 1e-3        # fp16 storage of the gradient


def _seg_model(size):
    from dose_prediction_b200 import networks
    from oracle import synth_ckpt
    tokens = (size // 16) ** 3
    man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("oar_transeg")]
    sd = synth_ckpt.make_state_dict(man, seed=1)
    model = networks.OARTranseg(1, 8, (size,) * 3, pos_embed="perceptron")
    model.load_state_dict(sd, strict=True)
    return model.to(DEV).train(), sd


def test_seg_training_step_matches_oracle_32():
    """One Transeg.training_step (DiceCE) + AdamW at 32^3, batch 2, vs the oracle's autograd; then the loss goes down."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import SegTrainer
    from oracle import torch_ref
    model, sd = _seg_model(32)
    vol = synth.make_batch(2, 32, seed=1234)
    label = synth.oar_labels(vol["oars"])
    loss_ref, grads_ref, new_ref, logits_ref = torch_ref.oar_transeg_train_step(sd, vol["ct"], label)
    tr = SegTrainer(model, 2, 32)
    ct, lab = vol["ct"].to(DEV), label.to(DEV)
    loss = tr.step(ct, lab)
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    assert _rel(tr.logits(), logits_ref) < 1e-2
    g = tr.grads()
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked = 0
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:
            assert float(g[n].norm()) < 1e-3 * gmax, n
            continue
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), ref.flatten().double(), dim=0))
        assert cos > 0.99, (n, cos)
        assert abs(float(g[n].norm()) / float(ref.norm()) - 1.0) < 0.08, n
        checked += 1
    assert checked > 100
    losses = [float(loss)] + [float(tr.step(ct, lab)) for _ in range(7)]
    print("seg losses", losses)
    assert all(b < a for a, b in zip(losses, losses[1:])) and losses[-1] < losses[0] - 0.1, losses     # lr 1e-4: slow but monotone


def test_training_step_matches_oracle_64():
    """64^3, batch 1: several W tiles / row blocks per wgrad launch and 64 ViT tokens; loss and the heavy layers' gradients."""
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    from oracle import torch_ref
    model, sd = _dose_model(64)
    vol = synth.make_batch(1, 64, seed=77)
    loss_ref, grads_ref, _, outs_ref = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"])
    tr = DoseTrainer(model, 1, 64)
    loss = tr.forward_backward(vol["dose_input"].to(DEV), vol["gt"].to(DEV))
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    assert _rel(tr.outputs()[1][0], outs_ref[1][0]) < 1e-2
    g = tr.grads()
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
        assert cos > 0.99, (n, cos)
        assert abs(float(g[n].norm()) / float(grads_ref[n].norm()) - 1.0) < 0.06, n


def test_old_transeg_training_step_matches_oracle_32():
    """OldModels TRANSEG (train_light_transeg.py mode_model=1; BatchNorm in both decoder branches, bare 1^3 conv)."""
    from dose_prediction_b200 import networks, synth
    from dose_prediction_b200.training import SegTrainer
    from oracle import synth_ckpt, torch_ref
    man = [(k, ([1, 8, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("transeg_old_96")]
    sd = synth_ckpt.make_state_dict(man, seed=2)
    model = networks.TRANSEG(1, 8, (32,) * 3, pos_embed="perceptron")
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train()
    vol = synth.make_batch(2, 32, seed=1234)
    label = synth.oar_labels(vol["oars"])
    loss_ref, grads_ref, new_ref, logits_ref = torch_ref.oar_transeg_train_step(sd, vol["ct"], label, old=True)
    tr = SegTrainer(model, 2, 32)
    loss = tr.step(vol["ct"].to(DEV), label.to(DEV))
    torch.cuda.synchronize()
    tr.P.check_device_errors()
    assert abs(float(loss) - float(loss_ref)) <= 2e-3 * abs(float(loss_ref))
    assert _rel(tr.logits(), logits_ref) < 1e-2
    g = tr.grads()
    gmax = max(float(v.norm()) for v in grads_ref.values())
    checked = 0
    for n, ref in grads_ref.items():
        if float(ref.norm()) < 1e-4 * gmax:
            continue
        cos = float(F.cosine_similarity(g[n].flatten().double().cpu(), ref.flatten().double(), dim=0))
        assert cos > 0.99, (n, cos)
        checked += 1
    assert checked > 60
    bn = "decoder2.conv_block.cov_.conv_3.conv.1."
    after = model.state_dict()
    assert _rel(after[bn + "running_mean"], new_ref[bn + "running_mean"]) < 1e-3
    assert _rel(after[bn + "running_var"], new_ref[bn + "running_var"]) < 1e-3
