#!/usr/bin/env python
"""Does the row pitch of K-major operands matter?  Times the tcgen05 GEMM on shapes of the step and of the LBS blend
product with the natural pitch (K floats) and with rows padded to 128 bytes (32 floats)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import functional as SF
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_gemm import t_us  # noqa: E402  (runs its table too when imported as a script -- guarded below)

def run(M, N, K, padA, padB, ldc=None):
    ldA = (K + padA - 1) // padA * padA
    ldB = (K + padB - 1) // padB * padB
    A = torch.randn(M, ldA, device="cuda"); B = torch.randn(N, ldB, device="cuda")
    ldc = ldc or (N + 3) // 4 * 4
    out = torch.empty(M, ldc, device="cuda")
    fn = lambda: SF.gemm(A, B, a_strides=(ldA, 1), b_strides=(ldB, 1), m=M, n=N, k=K, precision="tf32", out=out[:, :N], prerounded=True)
    us = t_us(fn, it=10)
    print(f"M={M:6d} N={N:5d} K={K:5d} ldA={ldA:5d} ldB={ldB:5d}: {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s")

for (M, N, K) in [(8192, 2334, 444), (2016, 1536, 784), (2016, 588, 784), (2016, 392, 588), (4032, 512, 784), (2016, 784, 392), (32768, 1536, 196), (8192, 4096, 1024)]:
    run(M, N, K, 4, 4)
    run(M, N, K, 32, 32)
run(8192, 2334, 444, 4, 4, ldc=2336)
run(8192, 2334, 444, 32, 32, ldc=2336)
run(8192, 2334, 444, 32, 32, ldc=2368)
