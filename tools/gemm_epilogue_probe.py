#!/usr/bin/env python
"""What an epilogue costs: the feed-forward GEMM shapes of the step with no epilogue, bias, bias + exact-erf GELU (which also
stores the pre-activation) and dGELU, stand-alone from a CUDA graph."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from scat_b200 import functional as SF
from bench_gemm import t_us

for (M, N, K) in [(2016, 588, 784), (2016, 294, 392), (2016, 392, 588), (2016, 196, 294), (4032, 588, 392), (4032, 294, 196)]:
    ld = (K + 7) // 8 * 8
    ldc = (N + 7) // 8 * 8
    A = torch.randn(M, ld, device="cuda"); B = torch.randn(N, ld, device="cuda"); bias = torch.randn(N, device="cuda")
    aux = torch.randn(M, ldc, device="cuda"); out = torch.empty(M, ldc, device="cuda")
    res = {}
    for epi in ("none", "bias", "bias_gelu", "dgelu", "bias_resid"):
        kw = dict(a_strides=(ld, 1), b_strides=(ld, 1), m=M, n=N, k=K, precision="tf32", out=out[:, :N], prerounded=True, epilogue=epi)
        if epi in ("bias", "bias_gelu", "bias_resid"): kw["bias"] = bias
        if epi in ("dgelu", "bias_resid"): kw["aux_in"] = aux[:, :N]
        res[epi] = t_us(lambda: SF.gemm(A, B, **kw), it=10)
    print(f"M={M} N={N} K={K}: " + "  ".join(f"{k} {v:6.1f} us" for k, v in res.items()))
