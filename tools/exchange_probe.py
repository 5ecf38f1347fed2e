#!/usr/bin/env python
"""What the gradient exchange costs and where: step time of one rank's train step (B=96, tf32, bf16 seam) issued
(a) as one call without exchange, (b) as its three phases in one graph without exchange (the phase joins alone),
(c) with the peer-memory exchange after the last kernel, (d) with the exchange hidden under the backward phases.
Run under torchrun with 2+ ranks (c, d need peers) or alone (a, b).  Optional argv[1] = file for rank 0's kernel timeline
of (d)."""
import os, random, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from scat_b200 import synth
from scat_b200.hand_net import EncoderTransformer
from scat_b200.train_step import HeadTrainStep

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 96
mean = torch.from_numpy(synth.make_mean_params("hand"))
opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)


def make(overlap):
    net = EncoderTransformer(opt, mean, precision="tf32", backbone=torch.nn.Identity())
    sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    ts = HeadTrainStep(net, B, x2_dtype="bf16", overlap_exchange=overlap)
    x2, mf, labels = (torch.from_numpy(a).cuda() for a in synth.make_head_inputs(B, rank))
    ts.load_inputs(x2.bfloat16(), mf, labels)
    return ts


def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


random.seed(0)
out = {}
ts = make(True)
ts.set_mask()
out["a_single_call_no_exchange"] = timed(lambda: ts.step(allreduce=False))
ts._warm(0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for ph in (0, 1, 2):
        ts._enqueue(0, ph)
out["b_three_phases_no_exchange"] = timed(g.replay)
if world > 1:
    out["d_exchange_hidden_under_backward"] = timed(lambda: ts.step(allreduce=True))
    if len(sys.argv) > 1 and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                ts.step(allreduce=True)
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Mem" not in e.name]
        evs.sort(key=lambda e: e.time_range.start)
        groups, cur = [], []
        for e in evs:
            if cur and e.time_range.start - max(x.time_range.end for x in cur) > 100:
                groups.append(cur); cur = []
            cur.append(e)
        groups.append(cur)
        gk = groups[-1]; t0 = gk[0].time_range.start
        with open(sys.argv[1], "w") as f:
            f.write(f"# rank 0 of {world}, exchange hidden under the backward: {len(gk)} kernels, span {max(e.time_range.end for e in gk) - t0:.1f} us\n")
            for e in gk:
                name = e.name.replace("scat::(anonymous namespace)::", "").replace("void ", "")
                f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {name[:90]}\n")
    else:
        for _ in range(3):
            ts.step(allreduce=True)
        torch.cuda.synchronize()
    ts.close()
    ts2 = make(False)
    ts2.set_mask()
    out["c_exchange_after_last_kernel"] = timed(lambda: ts2.step(allreduce=True))
    out["a2_single_call_no_exchange"] = timed(lambda: ts2.step(allreduce=False))
    ts2.close()
    dist.barrier()
if rank == 0:
    print("EXCHANGE_PROBE world=%d ms/step: %s" % (world, {k: round(v, 4) for k, v in out.items()}))
if world > 1:
    dist.destroy_process_group()
