#!/usr/bin/env python
"""Run the tensor-core conv front end (fwd, then mask_bwd + split + wgrad + dgrad) at config-2 size a few times: a small
target for ncu, and a quick timing.  usage: prof_conv.py [B] [tc|fp32] [fp32|bf16 seam]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth, functional as SF
from oracle import head_oracle  # positional encoding table only (test infrastructure, not on the product path)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
tc = (sys.argv[2] != "fp32") if len(sys.argv) > 2 else True
seam = sys.argv[3] if len(sys.argv) > 3 else "fp32"
x2, _, _ = synth.make_head_inputs(B, 3)
W = synth.make_head_weights(8)
cw = torch.from_numpy(W["conv1x1_channel_reduction.weight"]).cuda().view(21, 512)
mt = torch.from_numpy(W["mask_token"]).cuda().view(-1)
pe = head_oracle.positional_encoding(21, 784)[0].cuda()
idx = torch.tensor([3, 7, 11, 19], dtype=torch.int32, device="cuda")
x2d = torch.from_numpy(x2).cuda()
if seam == "bf16":
    x2d = x2d.bfloat16()
dtok = torch.randn(B, 21, 784, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    SF.conv_pe_mask_fwd(x2d, cw, pe, mt, idx, True, tc=tc)
    ev[1].record()
    SF.conv_bwd(dtok, x2d, cw, idx, tc=tc)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"iter {it}: fwd {ev[0].elapsed_time(ev[1])*1e3:.1f} us  bwd {ev[1].elapsed_time(ev[2])*1e3:.1f} us", flush=True)
