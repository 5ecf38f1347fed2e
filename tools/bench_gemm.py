#!/usr/bin/env python
"""Time the GEMM kernels on the shapes of the head (CUDA events, warm L2) and a few probes of fixed overhead."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import functional as SF

def t_us(fn, it=20):
    """Replay `it` launches from a CUDA graph so host launch cost (ctypes, tensor-map encode) is not timed."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * it) * 1e3

if __name__ == "__main__":
    shapes = [  # (M, N, K, layout)
        (2016, 1536, 784, "nt"), (2016, 784, 512, "nt"), (2016, 588, 784, "nt"), (2016, 392, 588, "nt"),
        (2016, 1536, 392, "nt"), (2016, 392, 512, "nt"), (2016, 296, 392, "nt"), (2016, 196, 296, "nt"),
        (2016, 1536, 196, "nt"), (2016, 196, 512, "nt"),
        (2016, 784, 1536, "nn"), (2016, 512, 784, "nn"), (2016, 784, 588, "nn"), (2016, 588, 392, "nn"),
        (1536, 784, 2016, "tn"), (784, 512, 2016, "tn"), (588, 784, 2016, "tn"), (392, 588, 2016, "tn"), (196, 512, 2016, "tn"),
        (2016, 392, 32, "nt"), (2016, 392, 128, "nt"), (128, 64, 32, "nt"), (128, 64, 2048, "nt"), (4032, 784, 1536, "nn"),
    ]
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    tot = 0.0
    for M, N, K, lay in shapes:
        if prec == "bf16":
            M, N, K = ((v + 7) // 8 * 8 for v in (M, N, K))
        dt = torch.bfloat16 if prec == "bf16" else torch.float32
        A = torch.randn(M, K, device="cuda").to(dt); B = torch.randn(N, K, device="cuda").to(dt)
        if lay == "nt": a, b, sa, sb = A, B, (K, 1), (K, 1)
        elif lay == "nn": a, b, sa, sb = A, B.t().contiguous(), (K, 1), (1, N)
        else: a, b, sa, sb = A.t().contiguous(), B.t().contiguous(), (1, M), (1, N)
        out = torch.empty(M, N, device="cuda")
        if prec == "bf16":
            us = t_us(lambda: SF.gemm_bf16(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, out=out))
        else:
            us = t_us(lambda: SF.gemm(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision=prec, out=out, prerounded=True))
        tot += us
        print(f"{prec} {lay} M={M:5d} N={N:5d} K={K:5d}: {us:7.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s  {(M*K+N*K)*A.element_size()/us/1e3+M*N*4/us/1e3:7.1f} GB/s(min traffic)")
    print("total", tot)
