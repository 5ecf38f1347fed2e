#!/usr/bin/env python
"""Kernel timeline of the config-4 token transformer forward (n = 128 tokens x dim 196, B = 256, inference): which kernels
the 0.8 ms are made of.  usage: config4_timeline.py [tf32|bf16]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth, functional as SF
from scat_b200.vision_transformer import Transformer
from scat_b200.hand_net import PositionalEncoding

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
dev = "cuda"
B4, n, dim, heads = 256, 128, 196, 8
Wt = synth.make_token_weights(dim, heads)
tr = Transformer(dim=dim, depth=3, heads=heads, dim_head=64, mlp_dim=2 * dim)
tr.load_state_dict({k[len("transformer."):]: torch.from_numpy(v) for k, v in Wt.items() if k.startswith("transformer.")})
tr = tr.to(dev)
pe = PositionalEncoding(dim, max_len=n).pe[0].to(dev)
tok = torch.from_numpy(synth.make_token_inputs(B4, n, dim, 5)).to(dev)
idx4 = torch.tensor(list(range(0, 25)), dtype=torch.int32, device=dev)
mt4 = torch.from_numpy(Wt["mask_token"]).to(dev).view(-1)
fn = lambda: SF.token_transformer(tr, tok, mask_token=mt4, pe=pe, mask_idx=idx4, precision=prec)
with torch.no_grad():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"# config 4 {prec}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per forward (eager issue, 20 back to back)")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); fn()
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
half = evs[len(evs) // 2:]
t0 = half[0].time_range.start
for e in half:
    name = e.name.replace("scat::(anonymous namespace)::", "").replace("void ", "")
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {name[:100]}")
