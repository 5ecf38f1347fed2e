#!/usr/bin/env python
"""Bring-up check of the tcgen05 GEMM: each operand-layout case runs in its own process under a timeout so a
faulting kernel cannot take the other cases (or the box) down.  Usage: python tools/check_gemm_tc.py [case]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "nt_small": dict(M=128, N=64, K=32, a="k", b="k"),
    "nt_k64": dict(M=128, N=64, K=64, a="k", b="k"),
    "nt_k256": dict(M=256, N=128, K=256, a="k", b="k"),
    "nt_ragged": dict(M=2016, N=1536, K=784, a="k", b="k"),
    "nt_n392": dict(M=2016, N=392, K=588, a="k", b="k"),
    "nn_small": dict(M=128, N=64, K=32, a="k", b="mn"),
    "nn_dgrad": dict(M=2016, N=784, K=512, a="k", b="mn"),
    "tn_small": dict(M=128, N=64, K=32, a="mn", b="mn"),
    "tn_wgrad": dict(M=1536, N=784, K=2016, a="mn", b="mn"),
    "tk_mix": dict(M=300, N=200, K=100, a="mn", b="k"),
}


def run_case(name):
    import numpy as np, torch
    from scat_b200 import functional as SF
    c = CASES[name]
    M, N, K = c["M"], c["N"], c["K"]
    g = np.random.Generator(np.random.PCG64(1))
    A = torch.from_numpy(g.standard_normal((M, K)).astype(np.float32))
    B = torch.from_numpy(g.standard_normal((N, K)).astype(np.float32))
    ref = A.double() @ B.double().t()
    Ad = A.cuda() if c["a"] == "k" else A.t().contiguous().cuda()      # MN-major: stored [K, M]
    Bd = B.cuda() if c["b"] == "k" else B.t().contiguous().cuda()
    a_str = (K, 1) if c["a"] == "k" else (1, M)
    b_str = (K, 1) if c["b"] == "k" else (1, N)
    out = SF.gemm(Ad, Bd, a_strides=a_str, b_strides=b_str, m=M, n=N, k=K, precision="tf32")
    torch.cuda.synchronize()
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    t = None
    if M >= 1024:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            SF.gemm(Ad, Bd, a_strides=a_str, b_strides=b_str, m=M, n=N, k=K, precision="tf32", out=out)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20 * 1e3
    tf = f" {t:.1f} us {2*M*N*K/t/1e6:.1f} TFLOP/s" if t else ""
    print(f"[gemm_tc] {name}: M={M} N={N} K={K} a={c['a']} b={c['b']} rel_err={err:.3e}{tf} {'OK' if err < 3e-3 else 'MISMATCH'}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                if r.returncode != 0:
                    tail += " | rc=%d %s" % (r.returncode, (r.stderr.strip().splitlines() or [""])[-1][:300])
                print(tail, flush=True)
            except subprocess.TimeoutExpired:
                print(f"[gemm_tc] {name}: TIMEOUT", flush=True)
