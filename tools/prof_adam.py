"""Target of the ncu capture of the fused Adam kernel (profiles/r1_ncu_adam_metrics.txt): the head's 3.8 M parameters,
eight rotating buffer sets so every launch streams its 106 MB through HBM rather than L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import _lib  # noqa: E402
from scat_b200._lib import check, ptr, stream_ptr  # noqa: E402

lib = _lib.load()
n = 3795648
sets = [[torch.randn(n, device="cuda") * 0.01, torch.randn(n, device="cuda"), torch.zeros(n, device="cuda"),
         torch.zeros(n, device="cuda")] for _ in range(8)]
for k in range(24):
    a = sets[k % 8]
    check(lib.scat_adam_step(ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), n, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1 + k // 8, None, None,
                             None, stream_ptr()), "scat_adam_step")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(80):
    a = sets[k % 8]
    check(lib.scat_adam_step(ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), n, 1e-4, 0.9, 0.999, 1e-8, 0.0, 4 + k // 8, None, None,
                             None, stream_ptr()), "scat_adam_step")
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 80 * 1e3
print(f"ok {float(sets[0][0].abs().mean()):.6f}  {us:.2f} us per launch, {28.0 * n / us / 1e3:.0f} GB/s algorithmic")
