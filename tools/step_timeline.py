#!/usr/bin/env python
"""Kernel timeline of ONE replay of the train step's CUDA graph (CUPTI through torch.profiler; nsys is not installed):
per kernel the stream, start offset and duration, so the critical chain, the side-stream overlap and the launch gaps can
be read off.  usage: step_timeline.py [precision] [seam] > gpurun_out/timeline.txt"""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from scat_b200 import synth
from scat_b200.hand_net import EncoderTransformer
from scat_b200.train_step import HeadTrainStep

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
seam = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = 96
mean = torch.from_numpy(synth.make_mean_params("hand"))
opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
net = EncoderTransformer(opt, mean, precision=precision, backbone=torch.nn.Identity())
sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
sd["positionalEncoding.pe"] = net.positionalEncoding.pe
net.load_state_dict(sd, strict=True)
net = net.cuda()
ts = HeadTrainStep(net, B, x2_dtype=seam)
x2, mf, labels = (torch.from_numpy(a).cuda() for a in synth.make_head_inputs(B, 0))
ts.load_inputs(x2.bfloat16() if seam == "bf16" else x2, mf, labels)
random.seed(0)
for _ in range(5):
    ts.set_mask(); ts.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ts.set_mask(); ts.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
# split into replays at gaps > 100 us
groups, cur = [], []
for e in evs:
    if cur and e.time_range.start - max(x.time_range.end for x in cur) > 100:
        groups.append(cur); cur = []
    cur.append(e)
groups.append(cur)
g = groups[-1]
t0 = g[0].time_range.start
streams = sorted({e.device_resource_id if hasattr(e, "device_resource_id") else 0 for e in g})
print(f"# {precision} GEMMs, {seam} seam, B={B}: last of 3 replays, {len(g)} kernels, span {max(e.time_range.end for e in g) - t0:.1f} us, streams {streams}")
print(f"{'start_us':>9s} {'dur_us':>7s} {'gap_us':>7s} strm  kernel")
last_end = {}
for e in g:
    st = getattr(e, "device_resource_id", 0)
    gap = e.time_range.start - last_end.get(st, t0)
    last_end[st] = e.time_range.end
    name = e.name.replace("scat::(anonymous namespace)::", "").replace("void ", "")
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f} {gap:7.1f} {streams.index(st):4d}  {name[:100]}")
busy = {}
for e in g:
    st = getattr(e, "device_resource_id", 0)
    busy[st] = busy.get(st, 0.0) + (e.time_range.end - e.time_range.start)
print("# busy time per stream (us):", {streams.index(k): round(v, 1) for k, v in busy.items()})
