"""GEMM probe for the CTA-pair kernel: the four ViT linears of one layer (M = 4096 tokens, hidden 768), each as a CUDA graph of
20 back-to-back launches (device time per launch without host launch gaps).  DP_GEMM_PAIR = 0 / 1 / 128 / 256 selects the kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dose_prediction_b200.engine import Plan


def main():
    dev = torch.device("cuda:0")
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    shapes = [("qkv", 2304, 768, None, "qkv"), ("proj", 768, 768, None, "resid"), ("fc1", 3072, 768, "gelu", "f16"), ("fc2", 768, 3072, None, "resid")]
    for name, N, K, act, kind in shapes:
        P = Plan(dev)
        A = (torch.randn(M, K, device=dev) / K ** 0.5).half()
        W = torch.randn(N, K, device=dev).half()
        bias = torch.randn(N, device=dev)
        P.keep += [A, W, bias]
        reps = 1 if os.environ.get('NCU') else 20
        for _ in range(reps):
            if kind == "qkv":
                heads, hd, T = 12, 64, 512
                q = P.zeros((M // T * heads, T, hd), torch.float16); k = P.zeros((M // T * heads, T, hd), torch.float16)
                vt = P.zeros((M // T * heads, hd, T), torch.float16)
                P.gemm(A, W, M, N, K, qkv=(heads, hd, T, q, k, vt, hd ** -0.5))
            elif kind == "resid":
                x = P.zeros((M, N), torch.float32)
                P.gemm(A, W, M, N, K, bias=bias, resid=x, out_f32=x)
            else:
                h = P.zeros((M, N), torch.float16)
                P.gemm(A, W, M, N, K, bias=bias, act=act, out_f16=h)
        P.run(); torch.cuda.synchronize()
        if os.environ.get('NCU'):
            continue
        P.capture()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        P.replay(); torch.cuda.synchronize()
        e0.record(); P.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"{name:5s} M{M} N{N} K{K}: {us:7.1f} us/launch  {2.0 * M * N * K / us / 1e6:7.0f} TFLOP/s  (DP_GEMM_PAIR={os.environ.get('DP_GEMM_PAIR', '1')})", flush=True)
        P.check_device_errors()


if __name__ == "__main__":
    main()
