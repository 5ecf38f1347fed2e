#!/usr/bin/env python
"""Does tcgen05 kind::tf32 truncate or round its fp32 operands?  Positive inputs expose a truncation bias."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import functional as SF

g = np.random.Generator(np.random.PCG64(0))
M, N, K = 512, 512, 1024
A = torch.from_numpy(g.uniform(1.0, 2.0, (M, K)).astype(np.float32))
B = torch.from_numpy(g.uniform(1.0, 2.0, (N, K)).astype(np.float32))
ref = A.double() @ B.double().t()
out = SF.gemm(A.cuda(), B.cuda(), precision="tf32").double().cpu()
print("positive inputs: mean(out/ref-1) = %.3e   rms = %.3e" % (float((out / ref - 1).mean()), float((out / ref - 1).pow(2).mean().sqrt())))

def rn_tf32(x):
    i = x.view(torch.int32)
    i = (i + 0x00000FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)
def tr_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
ref_rn = rn_tf32(A).double() @ rn_tf32(B).double().t()
ref_tr = tr_tf32(A).double() @ tr_tf32(B).double().t()
print("vs round-to-nearest model: rms %.3e ; vs truncation model: rms %.3e" % (
    float((out / ref_rn - 1).pow(2).mean().sqrt()), float((out / ref_tr - 1).pow(2).mean().sqrt())))
A2 = torch.from_numpy(g.standard_normal((M, K)).astype(np.float32)); B2 = torch.from_numpy(g.standard_normal((N, K)).astype(np.float32))
ref2 = A2.double() @ B2.double().t()
o_raw = SF.gemm(A2.cuda(), B2.cuda(), precision="tf32").double().cpu()
o_rn = SF.gemm(rn_tf32(A2).cuda(), rn_tf32(B2).cuda(), precision="tf32").double().cpu()
den = ref2.abs().max()
print("N(0,1) inputs: raw err(max-norm) %.3e  L2 %.3e | pre-rounded RN: %.3e  L2 %.3e" % (
    float((o_raw - ref2).abs().max() / den), float((o_raw - ref2).norm() / ref2.norm()),
    float((o_rn - ref2).abs().max() / den), float((o_rn - ref2).norm() / ref2.norm())))
