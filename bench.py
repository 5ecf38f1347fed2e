#!/usr/bin/env python
"""Benchmark of the reg_transformer head train step (BASELINE.json metric: head train samples/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a kernels
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one synthetic batch (BASELINE config 2: B=96 per GPU, mask_rate 0.2, pl_reg,
iteration 3, heads 8): forward, path-length VJP, projection + losses, backward of all 35 head tensors + x2.grad +
main_feat.grad, and (N>1) the all-reduce of the flat gradient bucket.  The ResNet backbone is outside the step
(north_star: timed separately).  Prints ONE JSON line.

Headline configuration: GEMMs on tcgen05 kind::tf32 ("tf32" precision) with the backbone seam tensor x2 handed over as
bfloat16 (what a bf16 / autocast backbone emits; x2.grad goes back as bfloat16).  `value`
is device-resident, `e2e` moves every step's inputs from pinned host memory inside the timed region.  The other
combinations (fp32 seam = the reference's own tensor types, bf16 GEMMs, fp32 parity mode) are reported under "variants".
Extra records on rank 0 at N = 1 (skipped with --quick): roofline per kernel, CPU baselines, the reference arithmetic
run eagerly on the same GPU (the kernel to beat on the same box), BASELINE configs 1, 4 and 5, a parity check of the
headline configuration against the float64 oracle, a >= 1 s soak.
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
# algorithmic bytes per sample (SURVEY.md section 8d / DESIGN.md section 5); seam = bytes per x2 / x2.grad element
BYTES_STEP_FP32 = 4957484                                     # whole train step, fp32 seam, SURVEY.md section 8d
FLOPS_STEP = 687.0e6
X2_ELEMS = 512 * 784
TOK_BYTES = 21 * 784 * 4


def conv_bytes(seam_bytes):
    return dict(fwd=X2_ELEMS * seam_bytes + 2 * TOK_BYTES,            # read x2, write feat_visual + token matrix
                wgrad=X2_ELEMS * seam_bytes + TOK_BYTES,             # read x2, read d tokens
                dgrad=X2_ELEMS * seam_bytes + TOK_BYTES)             # write x2.grad, read d tokens


def step_bytes(seam_bytes):
    return BYTES_STEP_FP32 - 3 * X2_ELEMS * (4 - seam_bytes)          # x2 read twice, x2.grad written once


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
# reference arm / cpu baselines: the oracle port (torch CPU, all host threads) on the same workloads
# ---------------------------------------------------------------------------------------------------
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_reference_run(steps: int, warmup: int, batch: int = B_PER_GPU, max_seconds: float = 150.0):
    """BASELINE config 2 on the CPU: fwd + path-length VJP + losses + bwd (hand_net.py:355-398, train.py:165-206)."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    sample = (f"{r['steps']} steps of B={r['batch']} (fwd + path-length VJP + losses + bwd), fp32, torch CPU, {cpu_model()}; "
              f"ONE host process with {r['cores']} threads whatever --gpus is (the reference has no distributed code: "
              f"compare against the N = 1 line)")
    line = {
        "impl": "reference", "metric": "head_train_samples_per_s", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, "fp32", "fp32"),
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "host processes: 1 (independent of --gpus)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, precision, seam):
    return {"workload": "BASELINE config 2: reg_transformer head train step (fwd + pl VJP + proj/loss + bwd), "
                        "B=96 per GPU, mask_rate 0.2, pl_reg, iteration 3, vit_heads 8, ResNet-50 seam tensors "
                        "x2[B,512,28,28] + main_feat[B,1024]",
            "global_batch": B_PER_GPU * n_gpus, "batch_per_gpu": B_PER_GPU, "precision": precision, "x2_seam": seam,
            "parallelism": f"dp{n_gpus}", "backbone": "excluded (timed separately, north_star)",
            "l2": "no explicit flush: per-step working set (x2 + x2.grad 154-308 MB + 170 MB workspace) > 126 MB L2"}


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
    """(rows, N, K, layout) of every tensor-core GEMM LAUNCH of one train step (csrc/head.cu): with pl_reg the dgrad
    chain sweeps the real and the path-length cotangent stacked, 2M rows in one launch."""
    inner, dims = 64 * heads, synth.layer_dims()
    MR = 2 * M if pl_reg else M
    t = []
    for l, (d, hid, out) in enumerate(dims):
        last = l == len(dims) - 1
        t += [(M, 3 * inner, d, "nt"), (M, d, inner, "nt")]                 # qkv, out projection
        t += [(MR, inner, d, "nn"), (MR, d, 3 * inner, "nn")]               # their data gradients
        t += [(d, inner, M, "tn"), (3 * inner, d, M, "tn")]                 # their weight gradients
        if not last:                                                        # the last feed-forward stays fp32 FFMA
            t += [(M, hid, d, "nt"), (M, out, hid, "nt")]
            t += [(MR, hid, out, "nn"), (MR, d, hid, "nn")]
            t += [(out, hid, M, "tn"), (hid, d, M, "tn")]
    return t


def measure_tensor_peaks(dev):
    """cuBLAS through torch.matmul at 8192^3: the denominators for the tensor-bound kernels, measured in this run
    (MEASURED_PEAKS.json has the bf16 figure only; TF32 was assumed to be half of it in round 1)."""
    out = {}
    n = 8192
    for name, dt, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False)):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        best = min(time_kernel(lambda: torch.matmul(a, b), iters=5, warm=2) for _ in range(3))
        torch.backends.cuda.matmul.allow_tf32 = old
        out[name] = 2.0 * n ** 3 / best / 1e12
        del a, b
    return out


def measure_roofline(ts, net, lib, pk, dev, B, step_s, precision, seam, tpk):
    from scat_b200 import functional as SF
    from scat_b200._lib import ptr, check, stream_ptr
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))

    def dram(name):
        t = traffic.get(name)
        return None if not t else t["dram_read"] + t["dram_write"]

    # --- HBM-bound front end: the three persistent conv kernels, timed live with CUDA events.  Four x2 buffers rotate
    # (4 x 77 MB bf16 / 2 x 154 MB fp32 > 126 MB L2), so every launch streams from HBM ---
    seam_bytes = 2 if seam == "bf16" else 4
    cb = conv_bytes(seam_bytes)
    W = net.head_parameters()
    pe = net.positionalEncoding.pe[0]
    idx = ts.mask_dev[: ts.n_masked]
    cw = W[1].data.view(21, 512)
    mt = W[0].data.view(-1)
    dtok = torch.randn(B, 21, 784, device=dev)
    x2s = [ts.x2s[i % len(ts.x2s)] for i in range(2)]
    extra = [torch.empty_like(x2s[0]) for _ in range(2)] if seam == "bf16" else []
    for e in extra:
        e.copy_(x2s[0])
    x2s = x2s + extra
    tc = precision != "fp32"
    seam_id = 1 if seam == "bf16" else 0
    fv, tok = torch.empty(B, 21, 784, device=dev), torch.empty(B, 21, 784, device=dev)
    nsc = lib.scat_conv_tc_scratch_floats(B, 512, 784, 21) if tc else lib.scat_conv_bwd_scratch_floats(B, 512, 784, 21)
    scratch = torch.empty(nsc, device=dev)
    x2gs = [torch.empty_like(x2s[0]) for _ in range(len(x2s))]
    wg, mg = torch.empty(21, 512, device=dev), torch.empty(784, device=dev)
    turn = [0]

    def conv_fwd():
        x = x2s[turn[0] % len(x2s)]
        turn[0] += 1
        if tc:
            check(lib.scat_conv_pe_mask_fwd_tc(ptr(x), seam_id, ptr(cw), ptr(pe), ptr(mt), ptr(idx), ts.n_masked, 1, ptr(fv),
                                               ptr(tok), ptr(scratch), B, 512, 784, 21, stream_ptr()), "conv fwd")
        else:
            check(lib.scat_conv_pe_mask_fwd(ptr(x), ptr(cw), ptr(pe), ptr(mt), ptr(idx), ts.n_masked, 1, ptr(fv), ptr(tok), B,
                                            512, 784, 21, stream_ptr()), "conv fwd")

    def conv_bwd():
        k = turn[0] % len(x2s)
        turn[0] += 1
        if tc:
            check(lib.scat_conv_bwd_tc(ptr(dtok), ptr(x2s[k]), seam_id, ptr(cw), ptr(idx), ts.n_masked, ptr(x2gs[k]), ptr(wg),
                                       ptr(mg), ptr(scratch), B, 512, 784, 21, stream_ptr()), "conv bwd")
        else:
            check(lib.scat_conv_bwd(ptr(dtok), ptr(x2s[k]), ptr(cw), ptr(idx), ts.n_masked, ptr(x2gs[k]), ptr(wg), ptr(mg),
                                    ptr(scratch), B, 512, 784, 21, stream_ptr()), "conv bwd")
    t_fwd = graph_time(conv_fwd, it=4, reps=5)
    t_bwd = graph_time(conv_bwd, it=4, reps=5)
    how = f"persistent tcgen05 kernels, {seam} seam" if tc else "fp32 FFMA"
    kernels = [
        {"kernel": f"conv forward: weight prep + conv_fwd_tc_kernel (1x1 conv + PE + token mask; {how})", "bound": "hbm",
         "achieved": cb["fwd"] * B / t_fwd / 1e9, "peak": pk["hbm"], "unit": "GB/s", "us": t_fwd * 1e6,
         "traffic": dram(f"conv_fwd_tc_kernel/{seam}")},
        {"kernel": f"conv backward: weight prep + mask_bwd + split + conv_wgrad_tc_kernel + conv_dgrad_tc_kernel ({how})",
         "bound": "hbm", "achieved": (cb["dgrad"] + cb["wgrad"]) * B / t_bwd / 1e9, "peak": pk["hbm"], "unit": "GB/s",
         "us": t_bwd * 1e6,
         "traffic": None if dram(f"conv_dgrad_tc_kernel/{seam}") is None else
         dram(f"conv_dgrad_tc_kernel/{seam}") + dram(f"conv_wgrad_tc_kernel/{seam}")},
    ]
    del x2gs, extra
    # --- dominant kernel by time share: the tcgen05 GEMM, every launch shape of the step replayed from a CUDA graph ---
    # operands as the step holds them: TF32 path = fp32 already rounded by the producer, BF16 path = bf16 in HBM
    if precision != "fp32":
        bf = precision == "bf16"
        tot_t, tot_f, n_launch = 0.0, 0.0, 0
        pad = 8 if bf else 4
        for (M, N, K, lay) in step_gemm_table(B * 21):
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
            wgr = lay == "tn"                                            # weight gradients run split-K in the step
            if bf:
                t = graph_time(lambda: SF.gemm_bf16(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, out=out, split_k=wgr))
            else:
                t = graph_time(lambda: SF.gemm(a, b, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32", out=out,
                                               prerounded=True, split_k=wgr))
            tot_t += t
            tot_f += 2.0 * M * N * K
            n_launch += 1
        peak = tpk["bf16"] if bf else tpk["tf32"]
        kind = "kind::f16 (bf16 operands)" if bf else "kind::tf32"
        gemm = {"kernel": f"gemm_tc_kernel (tcgen05 {kind}, TMEM accumulators, TMA), {n_launch} launches/step",
                "bound": "tensor", "achieved": tot_f / tot_t / 1e12, "peak": peak, "unit": "TFLOP/s",
                "us": tot_t * 1e6, "traffic": dram("gemm_tc_kernel (qkv layer 0)"),
                "note": f"peak = cuBLAS {'bf16' if bf else 'TF32'} 8192^3 measured in this run (torch.matmul, best of 3x5); "
                        "every GEMM launch shape of the step timed alone from a CUDA graph (operands L2-warm, as inside the "
                        "step where each operand was just written by its producer) with the operand storage the step uses; "
                        "weight gradients split-K with reductions in L2, as in the step"}
        kernels.insert(0, gemm)
    # --- fused Adam over the flat head parameters: 28 B/element (p, g, m, v read; p, m, v written), HBM bound.  Buffers
    # (4 x 15 MB) fit L2, so eight independent sets are rotated to keep the traffic in HBM ---
    n_par = ts.flat_params.numel()
    sets = [[torch.randn(n_par, device=dev) * 0.01, torch.randn(n_par, device=dev), torch.zeros(n_par, device=dev),
             torch.zeros(n_par, device=dev)] for _ in range(8)]
    aturn = [0]

    def adam_once():
        a = sets[aturn[0] % 8]
        aturn[0] += 1
        check(lib.scat_adam_step(ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), n_par, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1 + aturn[0],
                                 None, None, None, stream_ptr()), "scat_adam_step")
    t_adam = time_kernel(adam_once, iters=40, warm=8)
    kernels.append({"kernel": "adam_kernel (fused optimiser step over the flat head parameters; outside the headline step)",
                    "bound": "hbm", "achieved": 28.0 * n_par / t_adam / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "us": t_adam * 1e6, "traffic": None})
    del sets
    for k in kernels:
        k["frac"] = k["achieved"] / k["peak"]
    dom = kernels[0]
    sb = step_bytes(seam_bytes)
    return {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
            "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": pk["src"],
            "share_of_step": dom["us"] * 1e-6 / step_s, "kernels": kernels,
            "tensor_peaks_measured_tflops": tpk,
            "step": {"algorithmic_bytes": sb * B, "hbm_floor_us": sb * B / (pk["hbm"] * 1e9) * 1e6,
                     "algorithmic_flops": FLOPS_STEP * B, "us": step_s * 1e6,
                     "frac_of_hbm_floor": sb * B / (pk["hbm"] * 1e9) / step_s,
                     "tflops": FLOPS_STEP * B / step_s / 1e12}}


# ---------------------------------------------------------------------------------------------------
# the reference arithmetic (oracle port) run eagerly on the SAME GPU: PyTorch eager -> cuBLAS / cuDNN, which is what
# the reference does on a B200 ("the kernel to beat on the same box")
# ---------------------------------------------------------------------------------------------------
def gpu_eager_baseline(dev, B, iters=10):
    from oracle import head_oracle
    W = synth.make_head_weights(8)
    P = {k: torch.from_numpy(v).to(dev) for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params("hand")).to(dev)
    x2, mf, labels = (torch.from_numpy(a).to(dev) for a in synth.make_head_inputs(B, 0))
    idx = torch.tensor([20, 17, 19, 11], device=dev)
    pe = head_oracle.positional_encoding(21, 784).to(dev)      # the reference holds it as a device buffer (hand_net.py:74)
    out = {}

    def step():
        return head_oracle.train_step(P, x2, mf, labels, mean, heads=8, iteration=3, pos_embed=True, mask_idx=idx,
                                      pl_reg=True, pe=pe)["loss"]

    def run(mode):
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = mode != "fp32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else contextlib.nullcontext()
        res = {}
        try:
            with ctx:
                t = time_kernel(step, iters=iters, warm=3)
                res["eager"] = {"samples_per_s": B / t, "ms_per_step": t * 1e3}
                try:                                   # the same eager program captured once and replayed as a CUDA graph
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for _ in range(3):
                            step()
                    torch.cuda.current_stream().wait_stream(s)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        step()
                    t = time_kernel(g.replay, iters=iters, warm=3)
                    res["cuda_graph"] = {"samples_per_s": B / t, "ms_per_step": t * 1e3}
                    del g
                except Exception as e:                 # capture of an autograd step is not always possible
                    torch.cuda.synchronize()
                    res["cuda_graph"] = {"unavailable": str(e).splitlines()[0][:160]}
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        return res

    for mode in ("fp32", "tf32_allowed", "bf16_autocast"):
        out[mode] = run(mode)
    out["what"] = ("oracle/head_oracle.train_step (the reference's ATen call sequence, hand_net.py:355-398 + train.py:165-206) "
                   f"on cuda:0, B={B}: PyTorch eager -> cuBLAS / cuDNN sm_100 kernels, CUDA events, {iters} iterations")
    return out


def parity_record(net_factory, precision, seam, dev):
    """The headline configuration against the float64 oracle at BASELINE config 2 size, in this run."""
    from oracle import head_oracle
    from scat_b200.train_step import HeadTrainStep
    B = B_PER_GPU
    W = synth.make_head_weights(8)
    x2, mf, labels = synth.make_head_inputs(B, 11)
    x2t = torch.from_numpy(x2)
    if seam == "bf16":
        x2t = x2t.bfloat16().float()                    # identical inputs on both sides: the values a bf16 backbone delivers
    net = net_factory(precision)
    ts = HeadTrainStep(net, B, use_graph=False, x2_dtype=seam)
    ts.load_inputs(x2t.to(dev).to(ts.x2s[0].dtype), torch.from_numpy(mf).to(dev), torch.from_numpy(labels).to(dev))
    random.seed(3)
    mask = ts.set_mask()
    ts.step()
    torch.cuda.synchronize()
    P = {k: torch.from_numpy(v).double() for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params("hand")).double()
    o = head_oracle.train_step(P, x2t.double(), torch.from_numpy(mf).double(), torch.from_numpy(labels).double(), mean, heads=8,
                               iteration=3, pos_embed=True, mask_idx=mask, pl_reg=True)

    def l2(a, b):
        a, b = a.double().cpu(), b.double()
        return float((a - b).norm() / b.norm())

    def mx(a, b):
        a, b = a.double().cpu(), b.double()
        return float((a - b).abs().max() / b.abs().max())
    named = dict(net.named_parameters())
    g_l2 = {k: l2(named[k].grad, o["grads"][k]) for k in W}
    g_mx = {k: mx(named[k].grad, o["grads"][k]) for k in W}
    g_l2["main_feat"], g_mx["main_feat"] = l2(ts.main_feat_grad, o["main_feat_grad"]), mx(ts.main_feat_grad, o["main_feat_grad"])
    rec = {"against": "oracle/head_oracle.train_step in float64 on identical inputs and weights, B=96",
           "pred_max_rel": mx(ts.pred, o["pred"]), "joint1_exact_zero": bool(torch.all(ts.pred[:, 6:9] == 0)),
           "feat_visual_rel_l2": l2(ts.feat_visual, o["feat_visual"]), "pl_rel_l2": l2(ts.pl, o["pl"]),
           "loss_rel": abs(float(ts.losses[0]) - float(o["loss"])) / abs(float(o["loss"])),
           "param_grad_rel_l2_worst": max(g_l2.values()), "param_grad_max_norm_worst": max(g_mx.values()),
           "x2_grad_rel_l2": l2(ts.x2_grad, o["x2_grad"]), "x2_grad_max_norm": mx(ts.x2_grad, o["x2_grad"]),
           "masked_pl_rows_exact_zero": bool(torch.all(ts.pl.view(B, 21, -1)[:, mask] == 0)),
           "tolerances": "north_star: outputs 1e-4 relative (tf32 path), gradients 1e-3 relative; x2.grad in bf16 storage "
                         "adds its own 2^-9 rounding"}
    del ts, net
    return rec


def configs_record(dev, pk, tpk):
    """BASELINE configs 1, 4 and 5 on this box: GPU numbers with their roofline, the CPU port beside them."""
    from types import SimpleNamespace
    from oracle import head_oracle, mano_oracle
    from scat_b200 import functional as SF
    from scat_b200 import _lib as L
    from scat_b200._lib import ptr, check, stream_ptr
    from scat_b200.hand_net import EncoderTransformer, PositionalEncoding
    from scat_b200.mano import ManoLayer
    from scat_b200.vision_transformer import Transformer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lib = L.load()
    rec = {}
    # ---- config 1: head inference, B = 8, no grad (hand_net.py:363-393) ----
    W = synth.make_head_weights(8)
    Pc = {k: torch.from_numpy(v) for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    x2, mf, _ = (torch.from_numpy(a) for a in synth.make_head_inputs(8, 0))
    mask1 = [20, 17, 19, 11]
    with torch.no_grad():
        def cpu1():
            head_oracle.head_forward(Pc, x2, mf, mean, heads=8, iteration=3, pos_embed=True, mask_idx=mask1, pl_reg=False)
        for _ in range(3):
            cpu1()
        ts_ = []
        for _ in range(15):
            t0 = time.perf_counter(); cpu1(); ts_.append(time.perf_counter() - t0)
    t_cpu1 = float(np.median(ts_))
    opt = SimpleNamespace(vit_heads=8, pl_reg=False, iteration=3, pos_embed=True, mask_rate=0.2)
    with contextlib.redirect_stdout(sys.stderr):
        net = EncoderTransformer(opt, mean, precision="tf32", backbone=torch.nn.Identity())
    sd = dict(Pc)
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    x2d, mfd = x2.to(dev), mf.to(dev)
    with torch.no_grad():
        t_gpu1 = time_kernel(lambda: net.forward_features(mfd, x2d, mask_idx=mask1), iters=50, warm=5)
    rec["config1_inference_B8"] = {
        "cpu": {"samples_per_s": 8 / t_cpu1, "ms": t_cpu1 * 1e3, "cores": cores, "kind": "port", "how": "median of 15, fp32, no grad"},
        "gpu": {"samples_per_s": 8 / t_gpu1, "ms": t_gpu1 * 1e3, "precision": "tf32",
                "how": "module call (forward_features) incl. host launch cost, 50 iterations"}}
    del net
    # ---- config 4: n = 128 tokens x dim 196, B = 256, inference (hand_net.py:193-203) ----
    B4, n, dim, heads = 256, 128, 196, 8
    Wt = synth.make_token_weights(dim, heads)
    tr = Transformer(dim=dim, depth=3, heads=heads, dim_head=64, mlp_dim=2 * dim)
    tr.load_state_dict({k[len("transformer."):]: torch.from_numpy(v) for k, v in Wt.items() if k.startswith("transformer.")})
    tr = tr.to(dev)
    pe = PositionalEncoding(dim, max_len=n).pe[0].to(dev)
    tok = torch.from_numpy(synth.make_token_inputs(B4, n, dim, 5)).to(dev)
    idx4 = torch.tensor(list(range(0, 25)), dtype=torch.int32, device=dev)
    mt4 = torch.from_numpy(Wt["mask_token"]).to(dev).view(-1)
    c4 = []
    with torch.no_grad():
        for prec in ("tf32", "bf16"):
            t = time_kernel(lambda: SF.token_transformer(tr, tok, mask_token=mt4, pe=pe, mask_idx=idx4, precision=prec), iters=10)
            peak = tpk["bf16"] if prec == "bf16" else tpk["tf32"]
            c4.append({"precision": prec, "ms": t * 1e3, "samples_per_s": B4 / t, "tflops": 297.4e6 * B4 / t / 1e12,
                       "frac_of_tensor_peak": 297.4e6 * B4 / t / 1e12 / peak})
    Pt = {k: torch.from_numpy(v) for k, v in Wt.items()}
    tok_c = tok[:32].cpu()
    with torch.no_grad():
        t0 = time.perf_counter()
        head_oracle.token_transformer_forward(Pt, tok_c, heads=heads, mask_idx=list(range(25)))
        t_c4 = time.perf_counter() - t0
    rec["config4_tokens_n128_B256"] = {"gpu": c4, "algorithmic_flops_per_sample": 297.4e6,
                                       "cpu": {"samples_per_s": 32 / t_c4, "cores": cores, "kind": "port",
                                               "sample": "one forward of 32 samples"}}
    del tr
    # ---- config 5: MANO LBS (fwd, bwd) and the autoregressive regressor, B = 1k .. 64k ----
    layer = ManoLayer(synth.make_mano_asset())                         # blend shapes on the tcgen05 GEMM (fp32 grade)
    layer_ffma = ManoLayer(synth.make_mano_asset(), precision="fp32")  # everything on CUDA cores
    wr, br = torch.from_numpy(W["regressor.weight"]).to(dev), torch.from_numpy(W["regressor.bias"]).to(dev)
    meand = mean.to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    sweep = []
    for Bs in (1024, 4096, 16384, 65536):
        rots = 0.5 * torch.randn(Bs, 3, device=dev, generator=g)
        poses = 0.3 * torch.randn(Bs, 45, device=dev, generator=g)
        betas = torch.randn(Bs, 10, device=dev, generator=g)
        out = torch.empty(Bs, 799, 3, device=dev)
        t_f = time_kernel(lambda: layer(rots, poses, betas, out=out), iters=10)
        t_ff = time_kernel(lambda: layer_ffma(rots, poses, betas, out=out), iters=5)
        gout = torch.randn(Bs, 799, 3, device=dev, generator=g)
        gr, gp, gb = torch.empty_like(rots), torch.empty_like(poses), torch.empty_like(betas)
        t_b = time_kernel(lambda: check(lib.scat_lbs_bwd(ptr(layer.derived), ptr(layer.hands_mean), ptr(rots), ptr(poses),
                                                          ptr(betas), ptr(gout), ptr(gr), ptr(gp), ptr(gb), Bs, stream_ptr()),
                                        "lbs_bwd"), iters=5)
        mfs = torch.relu(torch.randn(Bs, 1024, device=dev, generator=g))
        fo = 0.05 * torch.randn(Bs, 63, device=dev, generator=g)
        t_r = time_kernel(lambda: SF.regressor_fwd(mfs, fo, meand, wr, br, iteration=3, root_relative=True), iters=10)

        def row(t, nbytes):
            return {"us": t * 1e6, "samples_per_s": Bs / t, "gbs": nbytes * Bs / t / 1e9, "frac_hbm": nbytes * Bs / t / 1e9 / pk["hbm"]}
        sweep.append({"batch": Bs, "lbs_fwd": dict(row(t_f, 9820), how="setup + tcgen05 blend-shape GEMM (3xTF32) + skinning kernel"),
                      "lbs_fwd_ffma": dict(row(t_ff, 9820), gflops_fp32=1.19e6 * Bs / t_ff / 1e9),
                      "lbs_bwd": row(t_b, 9820), "regressor": row(t_r, 4624)})
        del rots, poses, betas, out, gout, mfs, fo
    r_, p_, b_ = synth.make_mano_inputs(1024, 0)
    asset = synth.make_mano_asset()
    t0 = time.perf_counter()
    mano_oracle.rot_pose_beta_to_mesh(r_, p_, b_, asset)
    t_lbs_cpu = time.perf_counter() - t0
    rec["config5_lbs_regressor_sweep"] = {
        "gpu": sweep, "bytes_per_sample": {"lbs": 9820, "regressor": 4624},
        "bound": "LBS is fp32-ALU bound on CUDA cores (1.19 MFLOP per 9.8 KB, SURVEY.md section 7); lbs_fwd moves the two "
                 "blend-shape contractions (0.65 MFLOP) to the tensor cores; the HBM fractions are reported as north_star asks",
        "cpu_lbs_B1024": {"samples_per_s": 1024 / t_lbs_cpu, "ms": t_lbs_cpu * 1e3, "kind": "port",
                          "how": "numpy restatement of mano.py:280-391, one call"}}
    return rec


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and with them its pinned staging buffers, first-touch) to the CPU cores NVML reports
    as local to its GPU: with 8 ranks on a two-socket host the end-to-end copies otherwise cross the socket link.
    Returns (description, original affinity) -- the CPU baseline restores the original mask."""
    try:
        orig = os.sched_getaffinity(0)
    except Exception:
        return "unavailable (no sched_getaffinity)", None
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = max(orig) + 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1} & set(orig)
        if cpus and cpus != set(orig):
            os.sched_setaffinity(0, cpus)
            return f"rank bound to the {len(cpus)} CPU cores local to GPU {bus} (of {len(orig)})", orig
        return f"GPU {bus}: all {len(orig)} allowed cores are local, no binding needed", orig
    except Exception as e:           # binding is an optimisation of the end-to-end leg only
        return f"not bound ({type(e).__name__}: {str(e)[:80]})", orig


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from scat_b200 import _lib, dp
    from scat_b200.hand_net import EncoderTransformer
    from scat_b200.train_step import HeadTrainStep
    from scat_b200.optim import HeadAdam
    from types import SimpleNamespace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_binding, orig_affinity = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    pk = peaks()
    precision, seam = args.precision, args.seam
    if precision == "fp32":
        seam = "fp32"
    seam_t = torch.bfloat16 if seam == "bf16" else torch.float32

    class Seam(torch.nn.Module):          # backbone excluded: the head consumes the seam tensors directly
        def forward(self, x):
            raise RuntimeError("backbone is outside the benchmarked step")

    opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    sd0 = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}

    def make_net(prec):
        with contextlib.redirect_stdout(sys.stderr):       # the module prints like the reference's; stdout carries ONE JSON line
            n_ = EncoderTransformer(opt, mean, precision=prec, backbone=Seam())
        sd = dict(sd0)
        sd["positionalEncoding.pe"] = n_.positionalEncoding.pe
        n_.load_state_dict(sd, strict=True)
        return n_.to(dev)

    net = make_net(precision)
    dp.broadcast_parameters(net.head_parameters(), 0)

    B = B_PER_GPU
    ts = HeadTrainStep(net, B, 1e5, 10.0, use_graph=not args.no_graph, input_slots=2, comm=args.comm,
                       phased=args.phased, x2_dtype=seam)
    # two distinct synthetic batches per rank, pinned on the host in the seam dtype (e2e) and resident on the device (value)
    host = []
    for s in range(2):
        x2, mf, lab = synth.make_head_inputs(B, 100 + 2 * rank + s)
        host.append((torch.from_numpy(x2).to(seam_t).pin_memory(), torch.from_numpy(mf).pin_memory(),
                     torch.from_numpy(lab).pin_memory()))
    for s in range(2):
        ts.load_inputs(*[t.to(dev) for t in host[s]], slot=s)
    random.seed(1234)          # every rank draws the same host mask sequence (same indices on all shards)

    def one_step(slot=0):
        ts.set_mask()
        return ts.step(slot=slot)

    # launches per step (counted by the library itself)
    c0 = lib.scat_launch_count()
    ts._enqueue()
    launches_per_step = int(lib.scat_launch_count() - c0)
    if ts.peer is not None:             # the peer-memory all-reduce kernel(s) captured with the step: one launch per part
        launches_per_step += 4 if ts.overlap_exchange else 1
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: the peer-memory exchange against NCCL on integer-valued data (exact in fp32), once, in the warm-up ----
    ar_check = None
    if world > 1 and ts.peer is not None:
        from scat_b200._lib import stream_ptr
        gen = torch.Generator(device=dev).manual_seed(777 + rank)
        vals = torch.randint(-64, 65, (ts.bucket.flat.numel(),), device=dev, generator=gen).float()
        keep = ts.bucket.flat.clone()
        ref = vals.clone()
        dist.all_reduce(ref)
        ts.bucket.flat.copy_(vals)
        barrier()
        ts.peer.enqueue(stream_ptr())
        barrier()
        same = bool(torch.equal(ts.bucket.flat, ref))
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ar_check = ("bitexact vs NCCL all_reduce on every rank" if flag.item() == 1.0 else "MISMATCH vs NCCL all_reduce") + \
                   f" ({ts.bucket.flat.numel()} integer-valued fp32 elements, {world} ranks)"
        ts.bucket.flat.copy_(keep)

    for i in range(max(args.warmup, 3)):
        one_step(i % 2)
    torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): inputs already in HBM, two input slots alternate ----
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        one_step(i % 2)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    if ts.peer is not None:
        ts.check_peers()
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(ts.losses[0].item())

    # ---- end to end: every step's inputs travel pinned host -> device inside the timed region, the step's
    # losses travel back.  Two input slots: the copy of batch i+1 (copy stream) overlaps the compute of batch i.
    loss_host = torch.empty(4).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_run(step_obj, batches, steps):
        for slot in range(2):                       # warm both slots' graphs
            step_obj.load_inputs(*batches[slot], slot=slot)
            step_obj.set_mask()
            step_obj.step(slot=slot)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        copy_stream.wait_event(e0)
        step_obj.load_inputs(*batches[0], slot=0, stream=copy_stream)
        for i in range(steps):
            if i + 1 < steps:
                step_obj.load_inputs(*batches[(i + 1) % 2], slot=(i + 1) % 2, stream=copy_stream)
            step_obj.set_mask()
            step_obj.step(slot=i % 2)
            loss_host.copy_(step_obj.losses, non_blocking=True)
        e1.record()
        barrier()
        t = max(e0.elapsed_time(e1) * 1e-3, 0.0)
        return t, time.perf_counter() - w0

    t_e2e, t_e2e_wall = e2e_run(ts, host, args.steps)
    h2d = sum(t.numel() * t.element_size() for t in host[0]) + 4 * ts.n_masked
    d2h = 16

    def allmax(vals):
        if world == 1:
            return vals
        tt = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.tolist()
    t_dev, t_e2e = allmax([t_dev, t_e2e])

    # ---- >= 1 s soak of the same step (the K timed steps above last ~20 ms) ----
    soak = None
    if not args.quick:
        n_soak = max(args.steps, int(1.2 / (t_dev / args.steps)))
        barrier()
        e0.record()
        for i in range(n_soak):
            one_step(i % 2)
        e1.record()
        barrier()
        (t_soak,) = allmax([e0.elapsed_time(e1) * 1e-3])
        soak = {"steps": n_soak, "seconds": t_soak, "value": B * world * n_soak / t_soak, "unit": "samples/s",
                "ms_per_step": t_soak / n_soak * 1e3}

    # the same step with the fused Adam update in the graph, after the all-reduce; reported
    # next to the headline, which stays forward + backward (+ all-reduce) as BASELINE.json defines it
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
    (t_opt,) = allmax([e0.elapsed_time(e1) * 1e-3])
    ts.flat_params.copy_(keep_w)
    with_opt = {"value": B * world * args.steps / t_opt, "unit": "samples/s", "ms_per_step": t_opt / args.steps * 1e3,
                "optimizer": "fused Adam (scat_adam_step), in the step's CUDA graph"}

    # ---- the other precision / seam combinations of BASELINE config 2, same run (N = 1) ----
    variants = None
    if world == 1 and not args.quick:
        variants = []
        combos = [("tf32", "fp32"), ("tf32", "bf16"), ("bf16", "fp32"), ("bf16", "bf16"), ("fp32", "fp32")]
        for prec_v, seam_v in combos:
            if (prec_v, seam_v) == (precision, seam):
                variants.append({"precision": prec_v, "x2_seam": seam_v, "value": B * args.steps / t_dev, "unit": "samples/s",
                                 "ms_per_step": t_dev / args.steps * 1e3, "headline": True})
                continue
            net2 = make_net(prec_v)
            ts2 = HeadTrainStep(net2, B, 1e5, 10.0, use_graph=not args.no_graph, input_slots=2, x2_dtype=seam_v)
            st2 = torch.bfloat16 if seam_v == "bf16" else torch.float32
            host2 = [(h[0].to(st2).pin_memory(), h[1], h[2]) for h in host] if st2 != seam_t else host
            for s in range(2):
                ts2.load_inputs(*[t.to(dev) for t in host2[s]], slot=s)
            for i in range(3):
                ts2.set_mask(); ts2.step(slot=i % 2)
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                ts2.set_mask(); ts2.step(slot=i % 2)
            e1.record()
            torch.cuda.synchronize()
            t2 = e0.elapsed_time(e1) * 1e-3
            v = {"precision": prec_v, "x2_seam": seam_v, "value": B * args.steps / t2, "unit": "samples/s",
                 "ms_per_step": t2 / args.steps * 1e3, "loss": float(ts2.losses[0].item())}
            if prec_v == precision and seam_v != seam:     # the end-to-end number with the other seam dtype
                te, _ = e2e_run(ts2, host2, args.steps)
                v["e2e"] = {"value": B * args.steps / te, "unit": "samples/s", "ms_per_step": te / args.steps * 1e3,
                            "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host2[0]) + 4 * ts2.n_masked}
            variants.append(v)
            del ts2, net2

    if rank == 0:
        total = B * world * args.steps
        line = {
            "metric": "head_train_samples_per_s", "value": total / t_dev, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[precision], "data": "synthetic",
            "config": workload_config(world, precision, seam),
            "e2e": {"value": total / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e / args.steps * 1e3, "wall_ms_per_step": t_e2e_wall / args.steps * 1e3,
                    "x2_seam": seam, "how": "HeadTrainStep.load_inputs from pinned host tensors on a copy stream (two input "
                                            "slots) + step(), losses copied back every step"},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "cuda_graph": not args.no_graph, "loss": loss_val,
            "allreduce": {"none": "none", "peer": ("NVLink peer-memory exchange in four parts hidden under the backward (gradients-ready hook), inside the step's CUDA graph"
                                   if getattr(ts, "overlap_exchange", False) else "one NVLink peer-memory kernel inside the step's CUDA graph"),
                          "nccl": "NCCL, three phases overlapped with the backward" if ts.phased else
                                  "NCCL after the step"}[ts.comm],
            "clocks": clocks, "with_optimizer": with_opt,
        }
        if ar_check is not None:
            line["allreduce_check"] = ar_check
        if soak is not None:
            line["soak"] = soak
        if variants is not None:
            line["variants"] = variants
        if not args.quick:
            tpk = measure_tensor_peaks(dev)
            line["roofline"] = measure_roofline(ts, net, lib, pk, dev, B, t_dev / args.steps, precision, seam, tpk)
        if orig_affinity is not None:
            os.sched_setaffinity(0, orig_affinity)      # the CPU baseline uses every host core again
        line["host_binding"] = host_binding
        cpu = cpu_reference_run(steps=8, warmup=1, max_seconds=30.0)
        line["cpu_baseline"] = {"value": cpu["samples_per_s"], "unit": "samples/s", "cores": cpu["cores"], "kind": "port",
                                "sample": f"{cpu['steps']} steps of B={cpu['batch']} on {cpu_model()} (oracle port, fp32, fp32 x2)"}
        if world == 1 and not args.quick:
            line["parity"] = parity_record(make_net, precision, seam, dev)
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev, B)
            line["configs"] = configs_record(dev, pk, tpk)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        ts.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SCAT_PRECISION", "tf32"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--seam", default=os.environ.get("SCAT_SEAM", "bf16"), choices=["fp32", "bf16"],
                    help="storage of the backbone seam tensors x2 / x2.grad")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + e2e + CPU baseline only")
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
