#!/usr/bin/env python
"""Benchmark of the reg_transformer head train step (BASELINE.json metric: head train samples/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a kernels
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one synthetic batch (BASELINE config 2: B=96 per GPU,
mask_rate 0.2, pl_reg, iteration 3, heads 8): forward, path-length VJP, projection + losses, backward of
all 35 head tensors + x2.grad + main_feat.grad, and (N>1) the NCCL all-reduce of the flat gradient bucket.
The ResNet backbone is outside the step (north_star: timed separately).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from scat_b200 import synth  # noqa: E402

B_PER_GPU = 96
# algorithmic bytes per sample (SURVEY.md section 8d / DESIGN.md section 5), fp32 I/O
BYTES_CONV_FWD = 512 * 784 * 4 + 2 * 21 * 784 * 4            # read x2, write feat_visual + token matrix
BYTES_CONV_DGRAD = 512 * 784 * 4 + 21 * 784 * 4              # write x2.grad, read d tokens
BYTES_CONV_WGRAD = 512 * 784 * 4 + 21 * 784 * 4              # read x2, read d tokens
BYTES_STEP = 4957484                                          # whole train step, SURVEY.md section 8d
FLOPS_STEP = 687.0e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sus=float(p["bf16_tflops_sustained"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (torch CPU, all host threads) on the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, batch: int = B_PER_GPU, max_seconds: float = 150.0):
    from oracle import head_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = synth.make_head_weights(8)
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    x2, mf, labels = (torch.from_numpy(a) for a in synth.make_head_inputs(batch, 0))
    random.seed(0)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        mask = synth.mask_indices(0.2)
        t0 = time.perf_counter()
        head_oracle.train_step(P, x2, mf, labels, mean, heads=8, iteration=3, pos_embed=True, mask_idx=mask, pl_reg=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > max_seconds and len(times) >= 3:
            break
    total = float(np.sum(times))
    return dict(samples_per_s=batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                steps=len(times), batch=batch)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    sample = f"{r['steps']} steps of B={r['batch']} (fwd + path-length VJP + losses + bwd), fp32, torch CPU, {cpu_model()}"
    line = {
        "impl": "reference", "metric": "head_train_samples_per_s", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "fp32"),
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, precision):
    return {"workload": "BASELINE config 2: reg_transformer head train step (fwd + pl VJP + proj/loss + bwd), "
                        "B=96 per GPU, mask_rate 0.2, pl_reg, iteration 3, vit_heads 8, ResNet-50 seam tensors "
                        "x2[B,512,28,28] + main_feat[B,1024]",
            "global_batch": B_PER_GPU * n_gpus, "batch_per_gpu": B_PER_GPU, "precision": precision,
            "parallelism": f"dp{n_gpus}", "backbone": "excluded (timed separately, north_star)",
            "l2": "no explicit flush: per-step working set (x2 154 MB + x2.grad 154 MB + workspace) > 126 MB L2"}


# ---------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def graph_time(fn, it=10, reps=5):
    """Average device time of fn() replayed from a CUDA graph (host launch cost excluded, L2 warm like in-step)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * it) * 1e-3


def step_gemm_table(M, heads=8, pl_reg=True):
    """(M, N, K, layout, launches per step) of every tensor-core GEMM of one train step (csrc/head.cu)."""
    inner, dims = 64 * heads, synth.layer_dims()
    sweeps = 2 if pl_reg else 1                      # dgrad chain runs for the path-length VJP and for the backward
    t = []
    for l, (d, hid, out) in enumerate(dims):
        last = l == len(dims) - 1
        t += [(M, 3 * inner, d, "nt", 1), (M, d, inner, "nt", 1)]
        t += [(M, inner, d, "nn", sweeps), (M, d, 3 * inner, "nn", sweeps)]
        t += [(d, inner, M, "tn", 1), (3 * inner, d, M, "tn", 1)]
        if not last:                                 # last feed-forward stays fp32 FFMA
            t += [(M, hid, d, "nt", 1), (M, out, hid, "nt", 1)]
            t += [(M, hid, out, "nn", sweeps), (M, d, hid, "nn", sweeps)]
            t += [(out, hid, M, "tn", 1), (hid, d, M, "tn", 1)]
    return t


def measure_roofline(ts, net, lib, pk, dev, B, step_s, precision):
    from scat_b200 import functional as SF
    from scat_b200._lib import ptr, check, stream_ptr
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))

    def dram(name):
        t = traffic.get(name)
        return None if not t else t["dram_read"] + t["dram_write"]

    # --- HBM-bound front end: conv kernels, timed live with CUDA events (inputs 154 MB > L2) ---
    x2d = ts.x2
    W = net.head_parameters()
    pe = net.positionalEncoding.pe[0]
    idx = ts.mask_dev[: ts.n_masked]
    cw = W[1].data.view(21, 512)
    dtok = torch.randn(B, 21, 784, device=dev)
    tc = precision != "fp32"        # TF32 / BF16 modes run the conv passes on the tensor cores (batched tcgen05 GEMM)
    t_fwd = time_kernel(lambda: SF.conv_pe_mask_fwd(x2d, cw, pe, W[0].data.view(-1), idx, True, tc=tc))
    nsc = lib.scat_conv_tc_scratch_floats(B, 512, 784, 21) if tc else lib.scat_conv_bwd_scratch_floats(B, 512, 784, 21)
    scratch = torch.empty(nsc, device=dev)
    x2g, wg, mg = torch.empty_like(x2d), torch.empty(21, 512, device=dev), torch.empty(784, device=dev)
    bwd_fn = lib.scat_conv_bwd_tc if tc else lib.scat_conv_bwd

    seam_id = 1 if x2d.dtype == torch.bfloat16 else 0

    def conv_bwd():
        if tc:
            check(bwd_fn(ptr(dtok), ptr(x2d), seam_id, ptr(cw), ptr(idx), ts.n_masked, ptr(x2g), ptr(wg), ptr(mg),
                         ptr(scratch), B, 512, 784, 21, stream_ptr()), "scat_conv_bwd_tc")
        else:
            check(bwd_fn(ptr(dtok), ptr(x2d), ptr(cw), ptr(idx), ts.n_masked, ptr(x2g), ptr(wg), ptr(mg),
                         ptr(scratch), B, 512, 784, 21, stream_ptr()), "scat_conv_bwd")
    t_bwd = time_kernel(conv_bwd)
    how = "tcgen05 kind::tf32 batched GEMM" if tc else "fp32 FFMA"
    kernels = [
        {"kernel": f"conv_pe_mask_fwd ({how})", "bound": "hbm", "achieved": BYTES_CONV_FWD * B / t_fwd / 1e9,
         "peak": pk["hbm"], "unit": "GB/s", "us": t_fwd * 1e6, "traffic": dram("conv_pe_mask_fwd")},
        {"kernel": f"conv_bwd (mask_bwd + dgrad + wgrad, {how})", "bound": "hbm",
         "achieved": (BYTES_CONV_DGRAD + BYTES_CONV_WGRAD) * B / t_bwd / 1e9, "peak": pk["hbm"], "unit": "GB/s",
         "us": t_bwd * 1e6,
         "traffic": None if dram("conv_dgrad") is None else dram("conv_dgrad") + dram("conv_wgrad_partial")},
    ]
    # --- dominant kernel by time share: the tcgen05 GEMM, every shape of the step replayed from a CUDA graph ---
    # operands as the step holds them: TF32 path = fp32 already rounded by the producer, BF16 path = bf16 in HBM
    gemm = None
    if precision != "fp32":
        bf = precision == "bf16"
        tot_t, tot_f, n_launch = 0.0, 0.0, 0
        pad = 8 if bf else 4
        for (M, N, K, lay, cnt) in step_gemm_table(B * 21):
            M, N, K = ((v + pad - 1) // pad * pad for v in (M, N, K))   # hidden 294 lives in padded (16-byte) rows in the step
            dt = torch.bfloat16 if bf else torch.float32
            A = torch.randn(M, K, device=dev).to(dt)
            Bm = torch.randn(N, K, device=dev).to(dt)
            if lay == "nt":
                a, b, sa, sb = A, Bm, (K, 1), (K, 1)
            elif lay == "nn":
                a, b, sa, sb = A, Bm.t().contiguous(), (K, 1), (1, N)
            else:
                a, b, sa, sb = A.t().contiguous(), Bm.t().contiguous(), (1, M), (1, N)
            out = torch.zeros(M, N, device=dev)
            wg = lay == "tn"                                             # weight gradients run split-K in the step
            if bf:
                t = graph_time(lambda: SF.gemm_bf16(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, out=out, split_k=wg))
            else:
                t = graph_time(lambda: SF.gemm(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32", out=out,
                                               prerounded=True, split_k=wg))
            tot_t += cnt * t
            tot_f += cnt * 2.0 * M * N * K
            n_launch += cnt
        peak = pk["bf16"] if bf else pk["bf16"] / 2.0    # TF32 dense = half of BF16 dense; BF16 burst peak is the measured one
        kind = "kind::f16 (bf16 operands)" if bf else "kind::tf32"
        gemm = {"kernel": f"gemm_tc_kernel (tcgen05 {kind}, TMEM accumulators, TMA), {n_launch} launches/step",
                "bound": "tensor", "achieved": tot_f / tot_t / 1e12, "peak": peak, "unit": "TFLOP/s",
                "us": tot_t * 1e6, "traffic": dram("gemm_tc_kernel<128> (qkv layer 0)"),
                "note": ("peak = measured cuBLAS bf16 burst" if bf else "peak = measured cuBLAS bf16 burst / 2 (TF32)") +
                        "; every GEMM shape of the step timed alone from a CUDA graph with the operand storage the step uses"
                        + "; weight gradients split-K with reductions in L2, as in the step"}
        kernels.insert(0, gemm)
    # --- fused Adam over the flat head parameters: 28 B/element (p, g, m, v read; p, m, v written), HBM bound.  Buffers
    # (4 x 15 MB) fit L2, so eight independent sets are rotated to keep the traffic in HBM ---
    n_par = ts.flat_params.numel()
    sets = [[torch.randn(n_par, device=dev) * 0.01, torch.randn(n_par, device=dev), torch.zeros(n_par, device=dev),
             torch.zeros(n_par, device=dev)] for _ in range(8)]
    turn = [0]

    def adam_once():
        a = sets[turn[0] % 8]
        turn[0] += 1
        check(lib.scat_adam_step(ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), n_par, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1 + turn[0],
                                 None, None, None, stream_ptr()), "scat_adam_step")
    t_adam = time_kernel(adam_once, iters=40, warm=8)
    kernels.append({"kernel": "adam_kernel (fused optimiser step over the flat head parameters; outside the headline step)",
                    "bound": "hbm", "achieved": 28.0 * n_par / t_adam / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "us": t_adam * 1e6, "traffic": None})
    del sets
    for k in kernels:
        k["frac"] = k["achieved"] / k["peak"]
    dom = kernels[0]
    return {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
            "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": pk["src"],
            "share_of_step": dom["us"] * 1e-6 / step_s, "kernels": kernels,
            "step": {"algorithmic_bytes": BYTES_STEP * B, "hbm_floor_us": BYTES_STEP * B / (pk["hbm"] * 1e9) * 1e6,
                     "algorithmic_flops": FLOPS_STEP * B, "us": step_s * 1e6,
                     "frac_of_hbm_floor": BYTES_STEP * B / (pk["hbm"] * 1e9) / step_s}}


def run_ours(args):
    import torch.distributed as dist
    from scat_b200 import _lib, dp
    from scat_b200 import functional as SF
    from scat_b200.hand_net import EncoderTransformer
    from scat_b200.train_step import HeadTrainStep
    from types import SimpleNamespace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    pk = peaks()

    class Seam(torch.nn.Module):          # backbone excluded: the head consumes the seam tensors directly
        def forward(self, x):
            raise RuntimeError("backbone is outside the benchmarked step")

    opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    with contextlib.redirect_stdout(sys.stderr):       # the module prints like the reference's; stdout carries ONE JSON line
        net = EncoderTransformer(opt, mean, precision=args.precision, backbone=Seam())
    sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    dp.broadcast_parameters(net.head_parameters(), 0)

    B = B_PER_GPU
    ts = HeadTrainStep(net, B, 1e5, 10.0, use_graph=not args.no_graph, input_slots=2, comm=args.comm,
                       phased=args.phased)
    # two distinct synthetic batches per rank, pinned on the host (e2e) and resident on the device (value)
    host = []
    for s in range(2):
        x2, mf, lab = synth.make_head_inputs(B, 100 + 2 * rank + s)
        host.append(tuple(torch.from_numpy(a).pin_memory() for a in (x2, mf, lab)))
    ts.load_inputs(*[t.to(dev) for t in host[0]])
    random.seed(1234)          # every rank draws the same host mask sequence (same indices on all shards)

    def one_step():
        ts.set_mask()
        return ts.step()

    # launches per step (counted by the library itself)
    c0 = lib.scat_launch_count()
    ts._enqueue()
    launches_per_step = int(lib.scat_launch_count() - c0)
    if ts.peer is not None:
        launches_per_step += 1          # the peer-memory all-reduce kernel captured behind the step
    torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    if ts.peer is not None and ts.peer.timed_out():
        raise RuntimeError("a rank never arrived at the gradient all-reduce (20 s bounded wait): results are invalid")
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(ts.losses[0].item())

    # ---- end to end: every step's inputs travel pinned host -> device inside the timed region, the step's
    # losses travel back.  Two input slots: the copy of batch i+1 (copy stream) overlaps the compute of batch i.
    loss_host = torch.empty(4).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    for slot in range(2):                       # warm both slots' graphs
        ts.load_inputs(*host[slot], slot=slot)
        ts.set_mask()
        ts.step(slot=slot)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    copy_stream.wait_event(e0)
    ts.load_inputs(*host[0], slot=0, stream=copy_stream)
    for i in range(args.steps):
        if i + 1 < args.steps:
            ts.load_inputs(*host[(i + 1) % 2], slot=(i + 1) % 2, stream=copy_stream)
        ts.set_mask()
        ts.step(slot=i % 2)
        loss_host.copy_(ts.losses, non_blocking=True)
    e1.record()
    barrier()
    t_e2e = max(e0.elapsed_time(e1) * 1e-3, 0.0)
    t_e2e_wall = time.perf_counter() - w0
    h2d = sum(t.numel() * t.element_size() for t in host[0]) + 4 * ts.n_masked
    d2h = 16

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = tt.tolist()

    # the same step with the fused Adam update in the graph, after the all-reduce; reported
    # next to the headline, which stays forward + backward (+ all-reduce) as BASELINE.json defines it
    from scat_b200.optim import HeadAdam
    keep_w = ts.flat_params.clone()
    adam = HeadAdam(net.head_parameters(), lr=1e-4)
    ts.attach_optimizer(adam)
    for _ in range(3):
        ts.set_mask(); ts.step(optimize=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        ts.set_mask(); ts.step(optimize=True)
    e1.record()
    barrier()
    t_opt = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t_opt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_opt = float(tt.item())
    ts.flat_params.copy_(keep_w)
    with_opt = {"value": B * world * args.steps / t_opt, "unit": "samples/s", "ms_per_step": t_opt / args.steps * 1e3,
                "optimizer": "fused Adam (scat_adam_step), in the step's CUDA graph"}

    # the other tensor-core path of BASELINE config 2 ("bf16 and TF32 paths"), device-resident, same run
    other = None
    other_prec = {"tf32": "bf16", "bf16": "tf32"}.get(args.precision)
    if other_prec is not None and world == 1:
        with contextlib.redirect_stdout(sys.stderr):
            net2 = EncoderTransformer(opt, mean, precision=other_prec, backbone=Seam())
        net2.load_state_dict(sd, strict=True)
        net2 = net2.to(dev)
        ts2 = HeadTrainStep(net2, B, 1e5, 10.0, use_graph=not args.no_graph)
        ts2.load_inputs(*[t.to(dev) for t in host[0]])
        for _ in range(3):
            ts2.set_mask(); ts2.step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ts2.set_mask(); ts2.step()
        e1.record()
        torch.cuda.synchronize()
        t2 = e0.elapsed_time(e1) * 1e-3
        other = {"precision": other_prec, "value": B * args.steps / t2, "unit": "samples/s",
                 "ms_per_step": t2 / args.steps * 1e3, "loss": float(ts2.losses[0].item())}
        del ts2, net2

    if rank == 0:
        roofline = measure_roofline(ts, net, lib, pk, dev, B, t_dev / args.steps, args.precision)
        cpu = cpu_reference_run(steps=8, warmup=1, max_seconds=30.0)
        total = B * world * args.steps
        line = {
            "metric": "head_train_samples_per_s", "value": total / t_dev, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": workload_config(world, args.precision),
            "e2e": {"value": total / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e / args.steps * 1e3, "wall_ms_per_step": t_e2e_wall / args.steps * 1e3},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "cuda_graph": not args.no_graph, "loss": loss_val,
            "allreduce": {"none": "none", "peer": "one NVLink peer-memory kernel inside the step's CUDA graph",
                          "nccl": "NCCL, three phases overlapped with the backward" if ts.phased else
                                  "NCCL after the step"}[ts.comm],
            "clocks": clocks, "roofline": roofline, "other_precision": other, "with_optimizer": with_opt,
            "cpu_baseline": {"value": cpu["samples_per_s"], "unit": "samples/s", "cores": cpu["cores"], "kind": "port",
                             "sample": f"{cpu['steps']} steps of B={cpu['batch']} on {cpu_model()} (oracle port, fp32)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SCAT_PRECISION", "tf32"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--comm", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1 gradient all-reduce: this library's NVLink peer-memory kernel (auto) or NCCL")
    ap.add_argument("--phased", action="store_true", help="--comm nccl: overlap the all-reduce in three phases")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
