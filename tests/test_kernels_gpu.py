"""Per-kernel parity on a real B200, every call going through the C ABI (engine.Plan emitters).

Floating-point kernels: the comparison target is a plain PyTorch fp32/fp64 evaluation of the same op on
the operands exactly as the kernel sees them (fp16-rounded where the kernel stores fp16), so the only
difference left is fp32 accumulation order; tolerances are written per test.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _plan():
    from dose_prediction_b200.engine import Plan
    return Plan(torch.device("cuda:0"))


def _raw_to_ncdhw(t):
    n, cb, d, h, w, _ = t.shape
    return t.permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, d, h, w)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _h(x):
    return x.half().float()


def _finish(P):
    torch.cuda.synchronize()
    P.check_device_errors()


def _act_from(P, x, lo):
    """fp32 NCDHW cuda tensor -> packed Act through dp_pack_ncdhw."""
    a = P.new_act(x.shape[0], x.shape[1], tuple(x.shape[2:]), lo=lo)
    P.keep.append(x)
    P.pack_input(x, a)
    return a


@pytest.mark.parametrize("cin,cout,k,dil,dims,mode", [
    (16, 16, 3, 1, (16, 16, 16), "p1"),
    (32, 16, 7, 1, (9, 20, 12), "p1"),          # ragged H/W tiles, 2 K-chunks, full 7x7 stage
    (16, 64, 7, 1, (8, 16, 16), "p1"),          # kh-group split stages; depth-pair mode (two output planes per tile, N = 128)
    (32, 64, 7, 1, (5, 16, 40), "p1"),          # depth-pair mode: odd D (half-empty last pair), 5 W tiles (partial group), 2 chunks
    (16, 64, 7, 1, (2, 8, 8), "p1"),            # depth-pair mode: a single pair, masked rows
    (16, 128, 7, 1, (4, 16, 8), "p1"),          # one kh row per stage
    (64, 64, 3, 1, (8, 16, 16), "p1"),
    (128, 256, 3, 1, (4, 8, 8), "p1"),          # H < 16 (masked rows), N = 256
    (16, 16, 3, 2, (12, 16, 16), "p1"),         # dilation 2
    (9, 16, 3, 1, (16, 16, 16), "p3"),          # padded input channels + 3-term operand split (split-half stacked kernel)
    (32, 16, 3, 1, (9, 20, 40), "p3"),          # split-half stacked: 4 chunks, ragged H, a partial W-tile group (5 tiles)
    (16, 16, 3, 1, (1, 16, 16), "p3"),          # split-half stacked: a single plane (item start == item end)
    (16, 16, 3, 1, (2, 8, 8), "p3"),            # split-half stacked: one tile per CTA, masked rows
    (32, 32, 3, 1, (8, 16, 8), "p2"),
    (32, 32, 3, 2, (8, 16, 16), "p3"),          # plain kernel, [W_hi | W_lo] folded into N (C_out <= 32)
    (16, 32, 5, 1, (6, 12, 20), "p3"),          # folded, k = 5, ragged tiles
    (48, 64, 3, 1, (8, 16, 16), "p3"),          # plain kernel, unfolded 3-term split
])
def test_conv3d_tc_matches_torch(cin, cout, k, dil, dims, mode):
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    N = 2
    x = torch.randn(N, cin, *dims, device=dev)
    w = torch.randn(cout, cin, k, k, k, device=dev) / (cin * k ** 3) ** 0.5
    bias = torch.randn(cout, device=dev) * 0.1
    P = _plan()
    a = _act_from(P, x, lo=(mode != "p1"))
    raw = P.get_raw(N, cout, dims)
    scale, shift = P.affine(cout, bias=bias)
    P.conv_tc([a], w, k, dil, mode, scale, shift, False, out_raw=raw)
    P.run()
    _finish(P)
    got = _raw_to_ncdhw(raw.t)
    xe, we = (_h(x), _h(w)) if mode == "p1" else (x, w)
    want = F.conv3d(xe.double().cpu(), we.double().cpu(), bias.double().cpu(), padding=dil * (k - 1) // 2, dilation=dil)
    tol = 2e-5 if mode != "p2" else 1e-3        # p2 keeps fp16-rounded weights
    if mode == "p2":
        want = F.conv3d(x.double().cpu(), _h(w).double().cpu(), bias.double().cpu(), padding=dil * (k - 1) // 2)
        tol = 2e-5
    assert _rel(got.cpu(), want) < tol
    # statistics accumulated by the epilogue from the fp32 accumulators
    st = raw.stats.view(N, cout, 2).cpu()
    assert torch.allclose(st[..., 0], want.sum((2, 3, 4)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[..., 1], (want ** 2).sum((2, 3, 4)), rtol=1e-4, atol=1e-2)


def test_conv3d_stack_split_half_many_items_per_cta_and_old_layout(monkeypatch):
    """more work items than SMs (the ring's plane counter carries over from item to item), hi/lo fp16 output with ReLU,
    and the slot-major folded layout (DP_STACK_SPLIT_HALF=0) giving the same numbers"""
    from dose_prediction_b200 import engine
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    N, C, dims = 5, 16, (16, 128, 128)        # 160 columns x 16 planes: the two-phase balanced distribution (a whole column per CTA,
    x = torch.randn(N, C, *dims, device=dev)  # then contiguous shares of the other 12) — smaller problems keep fixed segments
    w = torch.randn(C, C, 3, 3, 3, device=dev) / (C * 27) ** 0.5
    bias = torch.randn(C, device=dev) * 0.1
    outs = []
    for split in (True, False):
        monkeypatch.setattr(engine, "STACK_SPLIT_HALF", split)
        P = _plan()
        a = _act_from(P, x, lo=True)
        out = P.new_act(N, C, dims, lo=True)
        st = P.new_stats(N, C)
        P.conv_tc([a], w, 3, 1, "p3", *P.affine(C, bias=bias), True, out_act=out, stats=st)
        y = torch.zeros(N, C, *dims, device=dev)
        P.unpack(out, y)
        P.run()
        _finish(P)
        outs.append((y, st.view(N, C, 2).clone()))
    want = F.relu(F.conv3d(x.double(), w.double(), bias.double(), padding=1))
    for y, st in outs:
        assert _rel(y, want) < 2e-5
        assert torch.allclose(st[..., 0], want.sum((2, 3, 4)), rtol=1e-4, atol=1e-1)
    assert _rel(outs[0][0], outs[1][0]) < 1e-6


def test_conv3d_stack_7_balanced_distribution_matches_torch():
    """the 7^3 depth-stacked kernel under the two-phase balanced work distribution (160 columns > 148 SMs, shares that start
    and end inside a column: clipped ring windows at arbitrary depths), fp16 operands, folded eval BatchNorm + ReLU"""
    torch.manual_seed(8)
    dev = torch.device("cuda:0")
    N, C, dims = 5, 16, (16, 128, 128)
    x = torch.randn(N, C, *dims, device=dev)
    w = torch.randn(C, C, 7, 7, 7, device=dev) / (C * 343) ** 0.5
    bias = torch.randn(C, device=dev) * 0.1
    P = _plan()
    a = _act_from(P, x, lo=False)
    raw = P.get_raw(N, C, dims)
    P.conv_tc([a], w, 7, 1, "p1", *P.affine(C, bias=bias), False, out_raw=raw)
    P.run()
    _finish(P)
    want = F.conv3d(_h(x).double(), _h(w).double(), bias.double(), padding=3)
    got = _raw_to_ncdhw(raw.t)
    assert _rel(got, want) < 2e-5
    st = raw.stats.view(N, C, 2)
    assert torch.allclose(st[..., 0], want.sum((2, 3, 4)), rtol=1e-4, atol=1e-1)
    assert torch.allclose(st[..., 1], (want ** 2).sum((2, 3, 4)), rtol=1e-4, atol=1e-1)


def test_conv3d_tc_concat_parts_bn_fold_relu_fp16_out():
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    N, dims, C = 2, (8, 16, 16), 16
    xa, xb = torch.randn(N, C, *dims, device=dev), torch.randn(N, C, *dims, device=dev)
    w = torch.randn(C, 2 * C, 7, 7, 7, device=dev) / (2 * C * 343) ** 0.5
    bn = torch.nn.BatchNorm3d(C).to(dev).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.2, 0.2)
        bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
    bias = torch.randn(C, device=dev) * 0.1
    P = _plan()
    pa, pb = P.new_concat(N, [C, C], dims)
    P.keep += [xa, xb]
    P.pack_input(xa, pa)
    P.pack_input(xb, pb)
    out = P.new_act(N, C, dims, lo=True)
    st = P.new_stats(N, C)
    P.conv_tc([pa, pb], w, 7, 1, "p1", *P.affine(C, bias=bias, bn=bn), True, out_act=out, stats=st)
    y = torch.zeros(N, C, *dims, device=dev)
    P.unpack(out, y)
    P.run()
    _finish(P)
    with torch.no_grad():
        want = F.relu(F.batch_norm(F.conv3d(_h(torch.cat((xa, xb), 1)).double().cpu(), _h(w).double().cpu(),
                                            bias.double().cpu(), padding=3), bn.running_mean.double().cpu(),
                                   bn.running_var.double().cpu(), bn.weight.double().cpu(), bn.bias.double().cpu(),
                                   False, 0.0, bn.eps))
    assert _rel(y.cpu(), want) < 2e-5           # hi+lo output carries ~22 bits
    assert torch.allclose(st.view(N, C, 2)[..., 0].cpu(), want.sum((2, 3, 4)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("M,N,K,split", [(512, 768, 768, 1), (256, 768, 4096, 4), (8, 768, 32768, 8), (512, 2304, 768, 1)])
def test_gemm_tc_matches_torch(M, N, K, split):
    torch.manual_seed(2)
    dev = torch.device("cuda:0")
    A = (torch.randn(M, K, device=dev) / K ** 0.5).half()
    B = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    P = _plan()
    out = P.zeros((M, N), torch.float32)
    P.gemm_splitk(A, B, M, N, K, split, out, bias=bias)
    P.run()
    _finish(P)
    want = A.double().cpu() @ B.double().cpu().t() + bias.double().cpu()
    assert _rel(out.cpu(), want) < 1e-5


_PAIR_CHECK = r"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r)
from dose_prediction_b200.engine import Plan
dev = torch.device("cuda:0")
torch.manual_seed(12)
for M, N, K, act in [(4096, 3072, 768, "gelu"), (4096, 768, 3072, None), (384, 768, 768, None), (1000, 256, 64, None),
                     (4096, 2304, 768, None)]:
    A = (torch.randn(M, K, device=dev) / K ** 0.5).half()
    B = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev)
    x0 = x.clone()
    P = Plan(dev)
    if act:
        g16 = P.zeros((M, N), torch.float16)
        P.gemm(A, B, M, N, K, bias=bias, act=act, out_f16=g16)
    else:
        P.gemm(A, B, M, N, K, bias=bias, resid=x, out_f32=x)
    P.run()
    torch.cuda.synchronize()
    P.check_device_errors()
    lin = A.double() @ B.double().t() + bias.double()
    got, want, tol = (g16.double(), F.gelu(lin), 1e-3) if act else (x.double(), lin + x0.double(), 1e-5)
    err = float((got - want).norm() / want.norm())
    assert err < tol, (M, N, K, err)
print("pair ok")
"""


@pytest.mark.parametrize("mode", ["1", "128", "256"])
def test_gemm_cta_pair_kernel_matches_torch(mode):
    """the cta_group::2 kernel (256 x 256 / 256 x 128 tiles over CTA pairs; off by default, DP_GEMM_PAIR selects it, read once
    per process -> a subprocess): several tiles per pair (both TMEM accumulators in flight), long and short K, M not a multiple
    of the pair tile, bias + GELU / residual epilogues"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DP_GEMM_PAIR=mode)
    r = subprocess.run([sys.executable, "-c", _PAIR_CHECK % root], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pair ok" in r.stdout, r.stdout + r.stderr


def test_gemm_tc_epilogues_gelu_residual_fp16():
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    M, N, K = 200, 768, 3072
    A = (torch.randn(M, K, device=dev) / K ** 0.5).half()
    B = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev)
    x0 = x.clone()
    P = _plan()
    h16 = P.zeros((M, N), torch.float16)
    P.gemm(A, B, M, N, K, bias=bias, resid=x, out_f32=x, out_f16=h16)
    g16 = P.zeros((M, N), torch.float16)
    P.gemm(A, B, M, N, K, bias=bias, act="gelu", out_f16=g16)
    P.run()
    _finish(P)
    lin = A.double().cpu() @ B.double().cpu().t() + bias.double().cpu()
    assert _rel(x.cpu(), lin + x0.double().cpu()) < 1e-5
    assert _rel(h16.float().cpu(), lin + x0.double().cpu()) < 1e-3
    assert _rel(g16.float().cpu(), F.gelu(lin)) < 1e-3


@pytest.mark.parametrize("T,heads,hd", [(512, 6, 128), (8, 12, 64), (216, 12, 64), (27, 6, 128)])
def test_attention_pipeline_matches_torch(T, heads, hd):
    """qkv scatter GEMM -> QK^T -> softmax -> PV, i.e. monai SABlock.forward without the projections."""
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    Bn, hidden = 2, heads * hd
    M = Bn * T
    xin = torch.randn(M, hidden, device=dev).half()
    wqkv = (torch.randn(3 * hidden, hidden, device=dev) / hidden ** 0.5).half()
    P = _plan()
    q = P.zeros((Bn * heads, T, hd), torch.float16)
    k = P.zeros((Bn * heads, T, hd), torch.float16)
    Tp = (T + 7) // 8 * 8
    vt = P.zeros((Bn * heads, hd, Tp), torch.float16)
    s = P.zeros((Bn * heads, T, T), torch.float32)
    pr = P.zeros((Bn * heads, T, Tp), torch.float16)
    o = P.zeros((M, hidden), torch.float16)
    P.gemm(xin, wqkv, M, 3 * hidden, hidden, qkv=(heads, hd, T, q, k, vt, hd ** -0.5))
    P.gemm(q, k, T, T, hd, batch=Bn * heads, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=s)
    P.softmax(s, Bn * heads * T, T, pr)
    P.gemm(pr, vt, T, hd, Tp, batch=Bn * heads, a_batch_rows=T, b_batch_rows=hd, c_batch_stride=T * hidden,
           c_batch_period=heads, c_batch_stride2=hd, ldc=hidden, out_f16=o)
    P.run()
    _finish(P)
    qkv = (xin.double().cpu() @ wqkv.double().cpu().t()).reshape(Bn, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * hd ** -0.5, dim=-1)
    want = (att @ qkv[2]).permute(0, 2, 1, 3).reshape(M, hidden)
    assert _rel(o.float().cpu(), want) < 4e-3       # q, k, v, p rounded to fp16


@pytest.mark.parametrize("T,heads,hd", [(512, 6, 128), (512, 12, 64), (8, 12, 64), (216, 12, 64), (27, 6, 128), (1728, 3, 64)])
def test_fused_attention_matches_torch_and_the_unfused_pipeline(T, heads, hd):
    """dp_attention (scores / probabilities never leave the SM) == monai SABlock.forward without the projections."""
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    Bn, hidden = 2, heads * hd
    M = Bn * T
    xin = torch.randn(M, hidden, device=dev).half()
    wqkv = (torch.randn(3 * hidden, hidden, device=dev) / hidden ** 0.5).half()
    wqkv[:hidden] *= 3.0                             # sharper softmax: exercises the max subtraction
    P = _plan()
    q = P.zeros((Bn * heads, T, hd), torch.float16)
    k = P.zeros((Bn * heads, T, hd), torch.float16)
    Tp = (T + 7) // 8 * 8
    vt = P.zeros((Bn * heads, hd, Tp), torch.float16)
    o = P.zeros((M, hidden), torch.float16)
    o.fill_(float("nan"))
    P.gemm(xin, wqkv, M, 3 * hidden, hidden, qkv=(heads, hd, T, q, k, vt, hd ** -0.5))
    P.attention(q, k, vt, Bn, heads, T, hd, o)
    P.run()
    _finish(P)
    qkv = (xin.double().cpu() @ wqkv.double().cpu().t()).reshape(Bn, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * hd ** -0.5, dim=-1)
    want = (att @ qkv[2]).permute(0, 2, 1, 3).reshape(M, hidden)
    assert torch.isfinite(o).all()
    assert _rel(o.float().cpu(), want) < 4e-3       # q, k, v, p rounded to fp16


def test_norm_act_residual_and_chained_stats():
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    N, C, dims = 2, 16, (8, 16, 16)
    x = torch.randn(N, C, *dims, device=dev) * 2 + 3
    r = torch.randn(N, C, *dims, device=dev)
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    P = _plan()
    a = _act_from(P, x, lo=True)
    res = _act_from(P, r, lo=True)
    one, zero = P.affine(C)
    w = torch.zeros(C, C, 1, 1, 1, device=dev)
    w[torch.arange(C), torch.arange(C)] = 1.0
    raw = P.get_raw(N, C, dims)
    P.pointwise([(a, None, None)], w, None, out_raw=raw)          # identity 1x1 -> raw fp32 + stats
    out = P.new_act(N, C, dims, lo=True)
    st2 = P.new_stats(N, C)
    P.norm_act(raw, out, gamma=g, beta=b, act="relu", res=res, act_after_res="lrelu", stats_out=st2)
    y = torch.zeros(N, C, *dims, device=dev)
    P.unpack(out, y)
    P.run()
    _finish(P)
    want = F.leaky_relu(F.relu(F.instance_norm(x.double().cpu(), weight=g.double().cpu(), bias=b.double().cpu(), eps=1e-5))
                        + r.double().cpu(), 0.01)
    assert _rel(y.cpu(), want) < 1e-5
    assert torch.allclose(st2.view(N, C, 2)[..., 1].cpu(), (want ** 2).sum((2, 3, 4)), rtol=1e-4)


def test_pointwise_two_sources_with_norm_on_load():
    torch.manual_seed(6)
    dev = torch.device("cuda:0")
    N, C, dims = 2, 32, (8, 8, 16)
    x3, x7 = torch.randn(N, C, *dims, device=dev) + 1, torch.randn(N, C, *dims, device=dev) * 3
    w = torch.randn(C, 2 * C, 1, 1, 1, device=dev) / (2 * C) ** 0.5
    bias = torch.randn(C, device=dev)
    ident = torch.zeros(C, C, 1, 1, 1, device=dev)
    ident[torch.arange(C), torch.arange(C)] = 1.0
    P = _plan()
    a3, a7 = _act_from(P, x3, True), _act_from(P, x7, True)
    r3, r7 = P.get_raw(N, C, dims), P.get_raw(N, C, dims)
    P.pointwise([(a3, None, None)], ident, None, out_raw=r3)
    P.pointwise([(a7, None, None)], ident, None, out_raw=r7)
    out = P.get_raw(N, C, dims)
    P.pointwise([(r3, r3.stats, "mish"), (r7, r7.stats, "mish")], w, bias, out_raw=out)
    planar = P.zeros((N, 1) + dims, torch.float32)
    P.pointwise([(a3, None, None)], w[:1, :C], bias[:1], out_planar=planar)
    P.run()
    _finish(P)
    cat = torch.cat((F.mish(F.instance_norm(x3.double().cpu(), eps=1e-5)), F.mish(F.instance_norm(x7.double().cpu(), eps=1e-5))), 1)
    want = F.conv3d(cat, w.double().cpu(), bias.double().cpu())
    assert _rel(_raw_to_ncdhw(out.t).cpu(), want) < 1e-5
    assert _rel(planar.cpu(), F.conv3d(x3.double().cpu(), w[:1, :C].double().cpu(), bias[:1].double().cpu())) < 1e-5


@pytest.mark.parametrize("wide", [False, True])
def test_pointwise_wide_coarse_level_goes_through_the_tensor_cores(wide):
    """128 input channels (two 64-channel sources with IN + Mish on load) -> 64: engine.pointwise materialises the
    activated sources once and contracts with dp_conv3d_tc (k = 1); fp16 sources take fp16 operands, hi/lo sources
    the 3-term split."""
    torch.manual_seed(16)
    dev = torch.device("cuda:0")
    N, C, dims = 2, 64, (4, 8, 16)
    x3, x7 = torch.randn(N, C, *dims, device=dev) + 1, torch.randn(N, C, *dims, device=dev) * 3
    w = torch.randn(C, 2 * C, 1, 1, 1, device=dev) / (2 * C) ** 0.5
    bias = torch.randn(C, device=dev)
    P = _plan()
    a3, a7 = _act_from(P, x3, wide), _act_from(P, x7, wide)
    st3, st7 = P.new_stats(N, C), P.new_stats(N, C)
    y3, y7 = P.new_act(N, C, dims, lo=wide), P.new_act(N, C, dims, lo=wide)
    P.norm_act(a3, y3, identity=True, stats_out=st3)
    P.norm_act(a7, y7, identity=True, stats_out=st7)
    out = P.get_raw(N, C, dims)
    P.pointwise([(y3, st3, "mish"), (y7, st7, "mish")], w, bias, out_raw=out)
    P.run()
    _finish(P)
    assert any(st[2] in ("dp_conv3d_tc", "dp_pointwise_tc") for st in P.steps)
    e3, e7 = (x3, x7) if wide else (_h(x3), _h(x7))
    cat = torch.cat((F.mish(F.instance_norm(e3.double().cpu(), eps=1e-5)), F.mish(F.instance_norm(e7.double().cpu(), eps=1e-5))), 1)
    want = F.conv3d(cat, w.double().cpu(), bias.double().cpu())
    got = _raw_to_ncdhw(out.t).cpu()
    assert _rel(got, want) < (2e-5 if wide else 2e-3)
    st = out.stats.view(N, C, 2).cpu()
    assert torch.allclose(st[..., 0], got.double().sum((2, 3, 4)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("Cs,Co,dims,acts,lo,out_act", [
    ((16, 9), 16, (5, 6, 7), (None, None), True, False),          # res-block conv3 of net_B: cat(out_net_A, x), ragged tile
    ((16, 16), 16, (8, 16, 16), ("relu", "relu"), True, False),    # seg decoder 1^3: hi/lo sources, IN + ReLU on load
    ((32, 32), 32, (4, 8, 12), ("mish", "mish"), False, False),    # dose decoder 1^3: fp16 sources, IN + Mish on load
    ((64, 64), 64, (3, 5, 9), ("mish", "mish"), False, False),     # K = 128, C_out = 64 (shuffle-reduced statistics)
    ((16, 16), 16, (4, 8, 8), (None, None), False, True),          # OldModels TRANSEG: bare 1^3 conv, fp16 output
    ((8, 8, 8), 16, (4, 4, 8), ("lrelu", None, "relu"), True, False)])   # odd number of 8-channel blocks (zero block)
def test_pointwise_tc_matches_torch(Cs, Co, dims, acts, lo, out_act):
    """dp_pointwise_tc: tcgen05 1^3 conv with InstanceNorm + activation applied on load, vs fp64 torch."""
    torch.manual_seed(60 + sum(Cs) + Co)
    dev = torch.device("cuda:0")
    N = 2
    xs = [torch.randn(N, c, *dims, device=dev) * (1 + i) + 0.5 * i for i, c in enumerate(Cs)]
    w = torch.randn(Co, sum(Cs), 1, 1, 1, device=dev) / sum(Cs) ** 0.5
    bias = torch.randn(Co, device=dev)
    P = _plan()
    srcs, refs = [], []
    for x, act in zip(xs, acts):
        a = _act_from(P, x, lo)
        e = x if lo else _h(x)
        if act is None:
            srcs.append((a, None, None))
            refs.append(e.double().cpu())
        else:
            st = P.new_stats(N, x.shape[1])
            y = P.new_act(N, x.shape[1], dims, lo=lo)
            P.norm_act(a, y, identity=True, stats_out=st)
            srcs.append((y, st, act))
            f = {"relu": F.relu, "mish": F.mish, "lrelu": lambda t: F.leaky_relu(t, 0.01)}[act]
            refs.append(f(F.instance_norm(e.double().cpu(), eps=1e-5)))
    want = F.conv3d(torch.cat(refs, 1), w.double().cpu(), bias.double().cpu())
    if out_act:
        out = P.new_act(N, Co, dims, lo=True)
        P.pointwise(srcs, w, bias, out_act=out)
        y = torch.zeros(N, Co, *dims, device=dev)
        P.unpack(out, y)
    else:
        out = P.get_raw(N, Co, dims)
        P.pointwise(srcs, w, bias, out_raw=out)
    P.run()
    _finish(P)
    assert any(st[2] == "dp_pointwise_tc" for st in P.steps)
    got = (y if out_act else _raw_to_ncdhw(out.t)).cpu()
    assert _rel(got, want) < 2e-5
    if not out_act:
        st = out.stats.view(N, Co, 2).cpu()
        assert torch.allclose(st[..., 0], got.double().sum((2, 3, 4)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(st[..., 1], (got.double() ** 2).sum((2, 3, 4)), rtol=1e-4, atol=1e-2)


def test_pointwise_tc_two_stage_norm_on_load_and_stats_only_pass():
    """conv_3_1's 3^3 branch without its materialised output: a statistics-only norm pass over the raw conv output, then the
    1^3 conv applying relu(IN(raw)) and IN + Mish on load (PtcBlock::stats0) — vs fp64 torch and vs the materialised path"""
    torch.manual_seed(77)
    dev = torch.device("cuda:0")
    N, C, dims = 2, 16, (6, 10, 12)
    x3 = torch.randn(N, C, *dims, device=dev) * 2 + 0.3
    x7 = torch.randn(N, C, *dims, device=dev)
    w = torch.randn(C, 2 * C, 1, 1, 1, device=dev) / (2 * C) ** 0.5
    bias = torch.randn(C, device=dev)
    outs = []
    for fused in (True, False):
        P = _plan()
        raw3 = _raw_with_stats(P, x3)
        a7 = _act_from(P, x7, True)
        st7, y7 = P.new_stats(N, C), P.new_act(N, C, dims, lo=True)
        P.norm_act(a7, y7, identity=True, stats_out=st7)
        st3 = P.new_stats(N, C)
        if fused:
            P.norm_act(raw3, None, act="relu", stats_out=st3)
            src3 = (raw3, st3, "mish", raw3.stats, "relu")
        else:
            y3 = P.new_act(N, C, dims, lo=True)
            P.norm_act(raw3, y3, act="relu", stats_out=st3)
            src3 = (y3, st3, "mish")
        out = P.get_raw(N, C, dims)
        P.pointwise([src3, (y7, st7, "mish")], w, bias, out_raw=out)
        P.run()
        _finish(P)
        assert any(st[2] == "dp_pointwise_tc" for st in P.steps)
        outs.append(_raw_to_ncdhw(out.t).cpu())
    r3 = F.mish(F.instance_norm(F.relu(F.instance_norm(x3.double().cpu(), eps=1e-5)), eps=1e-5))
    r7 = F.mish(F.instance_norm(x7.double().cpu(), eps=1e-5))
    want = F.conv3d(torch.cat((r3, r7), 1), w.double().cpu(), bias.double().cpu())
    assert _rel(outs[0], want) < 2e-5 and _rel(outs[1], want) < 2e-5
    assert _rel(outs[0], outs[1]) < 2e-6


def _raw_with_stats(P, x):
    """fp32 NCDHW tensor -> Raw (c8 fp32) with its {sum, sumsq} instance statistics filled in"""
    N, C = x.shape[0], x.shape[1]
    raw = P.get_raw(N, C, tuple(x.shape[2:]))
    raw.t.copy_(x.view(N, C // 8, 8, *x.shape[2:]).permute(0, 1, 3, 4, 5, 2))
    raw.stats.view(N, C, 2)[..., 0] = x.double().sum((2, 3, 4))
    raw.stats.view(N, C, 2)[..., 1] = (x.double() ** 2).sum((2, 3, 4))
    pinned = raw.stats.clone()
    P.keep.append(pinned)
    P.add_py(lambda: raw.stats.copy_(pinned))          # Plan.run() zeroes the statistics arena at the start of a replay
    return raw


@pytest.mark.parametrize("dims", [(16, 16, 32), (9, 13, 40)])        # second: ragged tiles in every direction
def test_one_channel_res_block_direct_conv_and_closed_form_residual(dims):
    """monai UnetResBlock with in_channels = 1 (seg encoder1): conv1 as the exact fp32 direct conv (dp_conv3d_c1) and the
    residual branch norm3(conv3(x)) in closed form inside the norm pass (dp_norm_act_resx), vs fp64 torch."""
    from dose_prediction_b200 import networks
    torch.manual_seed(31)
    dev = torch.device("cuda:0")
    N = 2
    x = torch.randn(N, 1, *dims, device=dev) * 0.7 + 0.3
    blk = networks.UnetResBlock(1, 16).to(dev).eval()
    with torch.no_grad():
        for prm in blk.parameters():
            prm.mul_(3.0)
    P = _plan()
    a = _act_from(P, x, True)
    out = P.new_act(N, 16, dims, lo=True)
    networks._emit_res_block(P, blk, [a], out, networks.PREC_SEG, x_planar=x)
    y = torch.zeros(N, 16, *dims, device=dev)
    P.unpack(out, y)
    P.run()
    _finish(P)
    assert any(st[2] == "dp_conv3d_c1" for st in P.steps) and any(st[2] == "dp_norm_act_resx" for st in P.steps)
    xd = x.double().cpu()
    w1, w2, w3 = (c.conv.weight.double().cpu() for c in (blk.conv1, blk.conv2, blk.conv3))
    h = F.leaky_relu(F.instance_norm(F.conv3d(xd, w1, padding=1), eps=1e-5), 0.01)
    h = F.instance_norm(F.conv3d(h, w2, padding=1), eps=1e-5)
    want = F.leaky_relu(h + F.instance_norm(F.conv3d(xd, w3), eps=1e-5), 0.01)
    assert _rel(y.cpu(), want) < 2e-5


def test_patch_embedding_gather_equals_patchify(monkeypatch):
    """dp_gemm_patch_embed (A tiles gathered from the c8 activation by one 5-D TMA box with element strides) must give
    exactly what dp_patchify + dp_gemm_tc give: same K order, same accumulation order."""
    from dose_prediction_b200 import networks
    torch.manual_seed(41)
    dev = torch.device("cuda:0")
    N, S = 2, (32, 128, 128)
    vit = networks.ViT(in_channels=25, img_size=S, patch_size=(16, 16, 16), hidden_size=768, mlp_dim=3072, num_layers=1,
                       num_heads=6, pos_embed="perceptron").to(dev).eval()
    x = torch.randn(N, 25, *S, device=dev)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("DP_PATCH_GATHER", flag)
        P = _plan()
        a16, a9 = P.new_concat(N, [16, 9], S, lo=False)
        x16, x9 = x[:, :16].contiguous(), x[:, 16:].contiguous()
        P.keep += [x16, x9]
        P.pack_input(x16, a16)
        P.pack_input(x9, a9)
        z, _ = networks._emit_vit(P, vit, [a16, a9], N, S, ())
        P.run()
        _finish(P)
        assert any(st[2] == ("dp_gemm_patch_embed" if flag == "1" else "dp_patchify") for st in P.steps)
        outs.append(z.t.clone())
    assert torch.equal(outs[0], outs[1])
    # and against torch: tokens = LN(block(Linear(rearranged patches) + pos))
    tok = x.half().float().view(N, 25, 2, 16, 8, 16, 8, 16).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(N, 128, -1)
    lin = vit.patch_embedding.patch_embeddings[1]
    emb = tok.double().cpu() @ lin.weight.half().double().cpu().t() + lin.bias.double().cpu() + vit.patch_embedding.position_embeddings.double().cpu()
    assert emb.shape == (N, 128, 768)


def test_deconv2x_c8_and_token_inputs():
    torch.manual_seed(7)
    dev = torch.device("cuda:0")
    from dose_prediction_b200.engine import Tokens
    N, Ci, Co, dims = 2, 32, 16, (4, 8, 8)
    x = torch.randn(N, Ci, *dims, device=dev)
    w = torch.randn(Ci, Co, 2, 2, 2, device=dev) / Ci ** 0.5
    P = _plan()
    a = _act_from(P, x, lo=True)
    slot0, slot1 = P.new_concat(N, [Co, Co], tuple(2 * d for d in dims), lo=True)
    P.deconv2x(a, w, slot1)
    y = torch.zeros(N, Co, *[2 * d for d in dims], device=dev)
    P.unpack(slot1, y)
    tok = torch.randn(N, 8, 768, device=dev).half()           # [B, T=2*2*2, 768]
    wt = torch.randn(768, Co, 2, 2, 2, device=dev) / 768 ** 0.5
    o2 = P.new_act(N, Co, (4, 4, 4))
    P.deconv2x(Tokens(tok, (2, 2, 2)), wt, o2)
    y2 = torch.zeros(N, Co, 4, 4, 4, device=dev)
    P.unpack(o2, y2)
    P.run()
    _finish(P)
    assert _rel(y.cpu(), F.conv_transpose3d(x.double().cpu(), w.double().cpu(), stride=2)) < 1e-5
    feat = tok.float().view(N, 2, 2, 2, 768).permute(0, 4, 1, 2, 3)
    assert _rel(y2.cpu(), F.conv_transpose3d(feat.double().cpu(), wt.double().cpu(), stride=2)) < 1e-3


@pytest.mark.parametrize("Ci,Co,dims,lo,tc", [
    (32, 16, (4, 8, 8), True, True), (64, 32, (4, 6, 10), True, True), (64, 64, (3, 5, 8), False, True),
    (128, 64, (2, 4, 4), False, True), (32, 16, (4, 8, 8), True, False), (64, 32, (4, 6, 10), False, False)])
def test_deconv2x_tensor_core_and_constant_bank_paths(Ci, Co, dims, lo, tc, monkeypatch):
    """ConvTranspose3d k2 s2 of a c8 tensor: dp_deconv2x_tc (1^3 implicit GEMM + scatter epilogue) and dp_deconv2x_cw."""
    from dose_prediction_b200 import engine
    monkeypatch.setattr(engine, "DECONV_TC", tc)
    torch.manual_seed(17)
    dev = torch.device("cuda:0")
    N = 2
    x = torch.randn(N, Ci, *dims, device=dev)
    w = torch.randn(Ci, Co, 2, 2, 2, device=dev) / Ci ** 0.5
    P = _plan()
    a = _act_from(P, x, lo=lo)
    slot0, slot1 = P.new_concat(N, [Co, Co], tuple(2 * d for d in dims), lo=lo)
    P.deconv2x(a, w, slot1)
    y = torch.zeros(N, Co, *[2 * d for d in dims], device=dev)
    P.unpack(slot1, y)
    P.run()
    _finish(P)
    xin = x if lo else _h(x)
    want = F.conv_transpose3d(xin.double().cpu(), w.double().cpu(), stride=2)
    assert _rel(y.cpu(), want) < (1e-5 if lo else 1.5e-3)      # lo: ~22-bit operands; else fp16 weights / outputs


def test_upsample_direct_conv_layernorm_patchify():
    torch.manual_seed(8)
    dev = torch.device("cuda:0")
    N, C, dims = 2, 16, (4, 6, 8)
    x = torch.randn(N, C, *dims, device=dev)
    P = _plan()
    a = _act_from(P, x, lo=True)
    up = P.new_act(N, C, tuple(2 * d for d in dims), lo=True)
    P.upsample2x(a, up)
    yu = torch.zeros(N, C, *[2 * d for d in dims], device=dev)
    P.unpack(up, yu)
    w = torch.randn(32, C, 3, 3, 3, device=dev) / (C * 27) ** 0.5
    bias = torch.randn(32, device=dev)
    raw = P.get_raw(N, 32, (2, 3, 4))
    P.conv_direct(a, w, 3, 2, 1, *P.affine(32, bias=bias), False, out_raw=raw)
    rows = torch.randn(37, 768, device=dev) * 2 + 1
    g, b = torch.rand(768, device=dev) + 0.5, torch.randn(768, device=dev)
    ln = P.zeros((37, 768), torch.float32)
    P.layernorm(rows, g, b, 37, 768, out_f32=ln)
    xv = torch.randn(N, 9, 32, 32, 32, device=dev)
    av = _act_from(P, xv, lo=False)
    A = P.zeros((N * 8, 2 * 4096 * 8), torch.float16)
    P.patchify(av, 2, A)
    P.run()
    _finish(P)
    assert _rel(yu.cpu(), F.interpolate(x.double().cpu(), scale_factor=2, mode="trilinear", align_corners=True)) < 1e-5
    assert _rel(_raw_to_ncdhw(raw.t).cpu(), F.conv3d(x.double().cpu(), w.double().cpu(), bias.double().cpu(), stride=2, padding=1)) < 1e-5
    assert _rel(ln.cpu(), F.layer_norm(rows.double().cpu(), (768,), g.double().cpu(), b.double().cpu(), 1e-5)) < 1e-5
    xp = torch.zeros(N, 16, 32, 32, 32)
    xp[:, :9] = _h(xv).cpu()
    want = xp.view(N, 2, 8, 2, 16, 2, 16, 2, 16).permute(0, 3, 5, 7, 1, 4, 6, 8, 2).reshape(N * 8, -1)
    assert torch.equal(A.float().cpu(), want)      # pure data movement: bit exact


def test_handoff_is_bit_exact():
    from oracle import torch_ref
    torch.manual_seed(9)
    dev = torch.device("cuda:0")
    S = 40
    logits = torch.randn(1, 8, S, S, S, device=dev)
    logits[0, :, :4] = 0.0                              # ties -> first index (background) must win
    ptv, ct = torch.rand(1, 1, S, S, S, device=dev), torch.randn(1, 1, S, S, S, device=dev)
    P = _plan()
    out = P.new_act(1, 9, (S, S, S), lo=True)
    st = P.zeros((1, 9, S, S, S), torch.float32)
    P.handoff(logits, ptv, ct, out, st)
    y = torch.zeros(1, 9, S, S, S, device=dev)
    P.unpack(out, y)
    P.run()
    _finish(P)
    want = torch_ref.handoff(logits.cpu(), ptv.cpu(), ct.cpu())
    assert torch.equal(st.cpu(), want)
    assert torch.allclose(y.cpu(), want, rtol=0, atol=1e-6)
