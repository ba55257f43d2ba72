"""Device time of one ViT layer's attention: fused dp_attention vs dp_gemm_tc + dp_softmax + dp_gemm_tc (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dose_prediction_b200.engine import Plan  # noqa: E402


def build(fused, Bn, heads, hd, T, layers=8):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    P = Plan(dev)
    hidden = heads * hd
    q = (torch.randn(Bn * heads, T, hd, device=dev) * hd ** -0.25).half()
    k = (torch.randn(Bn * heads, T, hd, device=dev) * hd ** -0.25).half()
    Tp = (T + 7) // 8 * 8
    vt = torch.zeros(Bn * heads, hd, Tp, device=dev, dtype=torch.float16)
    vt[:, :, :T] = torch.randn(Bn * heads, hd, T, device=dev).half()
    o = P.zeros((Bn * T, hidden), torch.float16)
    P.keep += [q, k, vt]
    if not fused:
        s = P.zeros((Bn * heads, T, T), torch.float32)
        pr = P.zeros((Bn * heads, T, Tp), torch.float16)
    for _ in range(layers):
        if fused:
            P.attention(q, k, vt, Bn, heads, T, hd, o)
        else:
            P.gemm(q, k, T, T, hd, batch=Bn * heads, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=s)
            P.softmax(s, Bn * heads * T, T, pr)
            P.gemm(pr, vt, T, hd, Tp, batch=Bn * heads, a_batch_rows=T, b_batch_rows=hd, c_batch_stride=T * hidden,
                   c_batch_period=heads, c_batch_stride2=hd, ldc=hidden, out_f16=o)
    return P, o


def main():
    for (Bn, heads, hd, T) in [(8, 12, 64, 512), (8, 6, 128, 512), (4, 12, 64, 1728), (4, 12, 64, 216)]:
        res = {}
        for fused in (False, True):
            P, o = build(fused, Bn, heads, hd, T)
            for _ in range(3):
                P.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                P.run()
            e1.record()
            torch.cuda.synchronize()
            P.check_device_errors()
            res[fused] = (e0.elapsed_time(e1) / 80 * 1e3, o.float().clone())
        d = (res[True][1] - res[False][1]).norm() / res[False][1].norm()
        print(f"B{Bn} heads{heads} hd{hd} T{T}: unfused {res[False][0]:.1f} us/layer, fused {res[True][0]:.1f} us/layer, rel diff {d:.2e}")


if __name__ == "__main__":
    main()
