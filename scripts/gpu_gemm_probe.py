"""GEMM micro-probe: times a few dp_gemm_tc shapes with CUDA events (used under ncu as well)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dose_prediction_b200.engine import Plan

def main():
    dev = torch.device("cuda:0")
    P = Plan(dev)
    T, hd, bh = 512, 64, 96
    q = torch.randn(bh, T, hd, device=dev).half(); k = torch.randn(bh, T, hd, device=dev).half()
    s = P.zeros((bh, T, T), torch.float32)
    P.gemm(q, k, T, T, hd, batch=bh, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=s)
    A = torch.randn(4096, 768, device=dev).half(); W = torch.randn(3072, 768, device=dev).half()
    bias = torch.randn(3072, device=dev)
    h = P.zeros((4096, 3072), torch.float16)
    P.gemm(A, W, 4096, 3072, 768, bias=bias, act="gelu", out_f16=h)
    h2 = P.zeros((4096, 3072), torch.float16)
    P.gemm(A, W, 4096, 3072, 768, out_f16=h2)
    for _ in range(3):
        P.run()
    torch.cuda.synchronize()
    for n, l, ms, _ in P.profile_launches():
        print(f"{ms:8.4f} ms {n} {l}")
    P.check_device_errors()

if __name__ == "__main__":
    main()
