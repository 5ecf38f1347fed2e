#!/usr/bin/env python
"""Quick timing of the MANO LBS forward (tensor-core and FFMA forms) and backward at 64 k samples."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth
from scat_b200.mano import ManoLayer

g = torch.Generator(device="cuda").manual_seed(1)
Bs = 65536
rots = 0.5 * torch.randn(Bs, 3, device="cuda", generator=g); poses = 0.3 * torch.randn(Bs, 45, device="cuda", generator=g)
betas = torch.randn(Bs, 10, device="cuda", generator=g)
tc, ff = ManoLayer(synth.make_mano_asset()), ManoLayer(synth.make_mano_asset(), precision="fp32")


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    t_tc = timed(lambda: tc(rots, poses, betas)); t_ff = timed(lambda: ff(rots, poses, betas))
    d = (tc(rots[:4096], poses[:4096], betas[:4096]) - ff(rots[:4096], poses[:4096], betas[:4096])).abs().max().item()
print(f"LBS_QUICK B={Bs}: tensor-core fwd {t_tc:.3f} ms = {Bs / t_tc / 1e3:.1f} M samples/s; FFMA fwd {t_ff:.3f} ms = {Bs / t_ff / 1e3:.1f} M/s; max |tc - ffma| = {d:.2e} m")
