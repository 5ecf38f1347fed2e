#!/usr/bin/env python
"""Where one tensor-core GEMM launch spends its time: CTA (0,0,0) records clock64() at 8 milestones."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import _lib, functional as SF
from scat_b200._lib import ptr

lib = _lib.load()
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
names = ["entry", "setup+dep wait done", "first stage landed", "accumulator ready", "chunk0 tmem ld", "chunk0 transposed", "chunk0 stored", "all tiles done"]
shapes = [(128, 64, 32), (128, 128, 2048), (2016, 784, 512), (2016, 1536, 784), (2016, 588, 784), (2016, 392, 588), (4032, 784, 1536)]
# output-bound persistent case shaped like the conv dgrad: 2352 tiles of 128x128 with 2 k-blocks, MN-major operands
M, N, K = 512, 96 * 784, 64
At = torch.randn(K, M, device="cuda"); Bt = torch.randn(K, N, device="cuda"); out = torch.empty(M, N, device="cuda")
fn = lambda: SF.gemm(At, Bt, a_strides=(1, M), b_strides=(1, N), m=M, n=N, k=K, precision="tf32", out=out, prerounded=True)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
lib.scat_debug_gemm_timeline(ptr(buf)); fn(); torch.cuda.synchronize(); lib.scat_debug_gemm_timeline(None)
t = buf.cpu().tolist()
print(f"dgrad-like persistent M={M} N={N} K={K}: {e0.elapsed_time(e1)*1e3:.1f} us; " +
      "  ".join(f"{n}={x - t[0]}" for n, x in zip(names[1:], t[1:])) + "  (SM cycles)")
for prec in ("tf32", "bf16"):
    for (M, N, K) in shapes:
        dt = torch.bfloat16 if prec == "bf16" else torch.float32
        A = torch.randn(M, K, device="cuda").to(dt); B = torch.randn(N, K, device="cuda").to(dt)
        out = torch.empty(M, N, device="cuda")
        fn = (lambda: SF.gemm_bf16(A, B, out=out)) if prec == "bf16" else (lambda: SF.gemm(A, B, precision="tf32", out=out, prerounded=True))
        for _ in range(3): fn()
        torch.cuda.synchronize()
        lib.scat_debug_gemm_timeline(ptr(buf))
        fn(); torch.cuda.synchronize()
        lib.scat_debug_gemm_timeline(None)
        t = buf.cpu().tolist()
        rel = [(x - t[0]) for x in t]
        print(f"{prec} M={M} N={N} K={K}: " + "  ".join(f"{n}={r}" for n, r in zip(names[1:], rel[1:])) + "  (SM cycles)")
