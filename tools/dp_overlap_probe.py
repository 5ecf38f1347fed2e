"""Where does the data-parallel overhead go?  Run under torchrun (N >= 2).  Times, per step (CUDA events + host clock):
single-graph step without all-reduce, three phase graphs without all-reduce, phased + overlapped all-reduce,
single graph + all-reduce afterwards, and the all-reduce alone at the sizes the step uses."""
import os
import random
import sys
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import dp, synth  # noqa: E402
from scat_b200.hand_net import EncoderTransformer  # noqa: E402
from scat_b200.train_step import HeadTrainStep  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if "--high-priority" in sys.argv:
        o = dist.ProcessGroupNCCL.Options()
        o.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=o)
    else:
        dist.init_process_group("nccl", device_id=dev)
    opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    net = EncoderTransformer(opt, mean, precision="tf32", backbone=torch.nn.Identity())
    sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    dp.broadcast_parameters(net.head_parameters(), 0)
    B = 96
    comm = "nccl" if "--nccl" in sys.argv else "peer"
    ts = HeadTrainStep(net, B, 1e5, 10.0, comm=comm, phased=(comm == "nccl"))
    x2, mf, lab = synth.make_head_inputs(B, 100 + rank)
    ts.load_inputs(*[torch.from_numpy(a).to(dev) for a in (x2, mf, lab)])
    random.seed(1)
    ts.set_mask()
    ts.step(); ts.step(allreduce=False)
    torch.cuda.synchronize()
    steps = 300
    cur = torch.cuda.current_stream()

    def timed(name, fn):
        for _ in range(20):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        host = time.perf_counter() - w0
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3, host / steps * 1e6], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:48s} device {t[0].item():8.1f} us/step   host issue {t[1].item():8.1f} us/step", flush=True)

    def phases_only():
        for g in ts.phase_graphs[0]:
            g.replay()

    def phases_sync_ar():       # phases, each followed by its all-reduce on the same stream (no overlap, same pieces)
        parts = ((ts.split, None), (ts.split0, ts.split), (0, ts.split0))
        for ph, g in enumerate(ts.phase_graphs[0]):
            g.replay()
            ts.bucket.all_reduce(ts.pg, lo=parts[ph][0], hi=parts[ph][1])

    timed("single graph, no all-reduce", lambda: ts.step(allreduce=False))
    if comm == "peer":
        from scat_b200._lib import stream_ptr
        timed("single graph incl. peer all-reduce (product)", lambda: ts.step())
        n = ts.peer.n_pad
        for name, lo, hi in (("whole bucket", 0, n), ("tail", ts.split // 4 * 4, n),
                             ("layer 0", ts.split0 // 4 * 4, ts.split // 4 * 4), ("head", 0, ts.split0 // 4 * 4)):
            timed(f"peer all-reduce alone: {name} {(hi - lo) * 4 / 1e6:.2f} MB", lambda: ts.peer.enqueue(stream_ptr(), lo, hi))
        torch.cuda.synchronize()
        if rank == 0:
            print("timed out:", ts.peer.timed_out(), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    timed("three phases, overlapped all-reduce (product)", lambda: ts.step())
    def partial(which):
        parts = ((ts.split, None), (ts.split0, ts.split), (0, ts.split0))
        def fn():
            for ph, g in enumerate(ts.phase_graphs[0]):
                g.replay()
                if ph in which and ph < 2:
                    ts.comm_stream.wait_stream(cur)
                    with torch.cuda.stream(ts.comm_stream):
                        ts.bucket.all_reduce(ts.pg, lo=parts[ph][0], hi=parts[ph][1])
            if 2 in which:
                ts.bucket.all_reduce(ts.pg, hi=ts.split0)
            cur.wait_stream(ts.comm_stream)
        return fn
    if "--parts" in sys.argv:
        timed("phases + tail AR only (overlaps phase 1)", partial((0,)))
        timed("phases + layer-0 AR only (overlaps phase 2)", partial((1,)))
        timed("phases + head AR only (exposed)", partial((2,)))
        timed("phases + tail + layer-0 AR", partial((0, 1)))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        for i, g in enumerate(ts.phase_graphs[0]):
            g.replay(); e[i + 1].record()
        torch.cuda.synchronize()
        if rank == 0:
            print("phase durations (us):", [round(e[i].elapsed_time(e[i + 1]) * 1e3, 1) for i in range(3)], flush=True)
    if "--quick" in sys.argv:
        dist.destroy_process_group()
        return
    timed("three phase graphs, no all-reduce", phases_only)
    timed("three phases, all-reduce in stream order", phases_sync_ar)
    ts.phased = False
    timed("single graph, whole-bucket all-reduce after", lambda: ts.step())
    ts.phased = True
    n = ts.bucket.flat.numel()
    for name, lo, hi in (("whole bucket", 0, n), ("tail (layers 1,2,regressor)", ts.split, n),
                         ("layer 0", ts.split0, ts.split), ("mask token + conv weight", 0, ts.split0)):
        part = ts.bucket.flat[lo:hi]
        timed(f"all-reduce alone: {name} {part.numel() * 4 / 1e6:.2f} MB", lambda: dist.all_reduce(part))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
