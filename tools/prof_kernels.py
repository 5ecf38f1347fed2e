#!/usr/bin/env python
"""A short sequence of the step's representative kernels at config-2 size (one launch each per iteration): the
target of the `ncu --set full` capture whose summary is committed under profiles/."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import functional as SF

M, inner = 2016, 512
dev = "cuda"
a32 = torch.randn(M, 784, device=dev); w32 = torch.randn(3 * inner, 784, device=dev); o32 = torch.empty(M, 3 * inner, device=dev)
a16, w16 = a32.bfloat16(), w32.bfloat16()
d16 = torch.randn(2 * M, 3 * inner, device=dev).bfloat16(); wq16 = torch.randn(3 * inner, 784, device=dev).bfloat16()
dn = torch.empty(2 * M, 784, device=dev)
qkv = torch.randn(M, 3 * inner, device=dev); d_o = torch.randn(2 * M, inner, device=dev)
x = torch.randn(2 * M, 784, device=dev); g = torch.ones(784, device=dev); b = torch.zeros(784, device=dev)
for it in range(3):
    SF.gemm(a32, w32, precision="tf32", out=o32, prerounded=True)                     # qkv layer 0, TF32
    SF.gemm_bf16(a16, w16, out=o32)                                                    # qkv layer 0, BF16
    SF.gemm_bf16(d16, wq16, a_strides=(3 * inner, 1), b_strides=(1, 784), m=2 * M, n=784, k=3 * inner, out=dn)   # dNa (stacked dgrad)
    o, p = SF.attention_fwd(qkv, 96, 21, 8, tc=True)
    SF.attention_bwd(qkv, p, d_o[:M], 96, 21, 8, tc=True)
    y, mean, rstd = SF.layernorm_fwd(x[:M], g, b)
    SF.layernorm_bwd(x[:M], x[:M], g, mean, rstd, resid=x[:M])
torch.cuda.synchronize()
print("ok")
