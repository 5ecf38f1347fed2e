#!/usr/bin/env python
"""Runs the kernels that had no ncu evidence in round 1 a few times each (a small target for `ncu --set full`):
attention_fwd_tc128 (config 4 shape), the LBS forward (FFMA and tensor-core forms) and backward, the regressor kernels,
the fp32 FFMA GEMM at the last feed-forward's shapes, proj_loss, and the peer all-reduce kernel with one rank (its
NVLink traffic needs >= 2 GPUs: see profiles/README.md)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth, functional as SF, dp
from scat_b200._lib import stream_ptr
from scat_b200.mano import ManoLayer

g = torch.Generator(device="cuda").manual_seed(0)
B, n, heads = 256, 128, 8
qkv = torch.randn(B * n, 3 * 64 * heads, device="cuda", generator=g)
W = synth.make_head_weights(8)
wr, br = torch.from_numpy(W["regressor.weight"]).cuda(), torch.from_numpy(W["regressor.bias"]).cuda()
mean = torch.from_numpy(synth.make_mean_params("hand")).cuda()
Bs = 16384
rots = 0.5 * torch.randn(Bs, 3, device="cuda", generator=g); poses = 0.3 * torch.randn(Bs, 45, device="cuda", generator=g)
betas = torch.randn(Bs, 10, device="cuda", generator=g)
tc, ff = ManoLayer(synth.make_mano_asset()), ManoLayer(synth.make_mano_asset(), precision="fp32")
mf = torch.relu(torch.randn(Bs, 1024, device="cuda", generator=g)); fo = 0.05 * torch.randn(Bs, 63, device="cuda", generator=g)
a = torch.randn(2016, 196, device="cuda", generator=g); w1 = torch.randn(147, 196, device="cuda", generator=g)
pred = torch.randn(96, 66, device="cuda", generator=g); labels = torch.randn(96, 105, device="cuda", generator=g)
pm = dp.PeerMemory(3795200, torch.device("cuda", 0))
for it in range(3):
    SF.attention_fwd(qkv, B, n, heads, tc=True)
    tc(rots, poses, betas)
    ff(rots, poses, betas)
    r, p, b = rots.clone().requires_grad_(True), poses.clone().requires_grad_(True), betas.clone().requires_grad_(True)
    tc(r, p, b).sum().backward()
    SF.regressor_fwd(mf, fo, mean, wr, br, iteration=3, root_relative=True)
    SF.gemm(a, w1, precision="fp32")
    SF.proj_loss(pred, labels, None)
    pm.enqueue(stream_ptr())
    torch.cuda.synchronize()
pm.close()
print("prof_misc done")
