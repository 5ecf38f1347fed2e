"""Throughput of the device evaluation metrics (Procrustes alignment + PCK counts + MPJPE + acceleration error) next to
the oracle port of the reference's functions on the host.  Prints one JSON line (profiles/r1_eval_metrics.json)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import eval_oracle  # noqa: E402
from scat_b200 import eval_metrics as EM  # noqa: E402

B = 65536
rng = np.random.default_rng(5)
gt = (rng.standard_normal((B, 21, 3)) * 0.04).astype(np.float32)
pred = (gt * 1.1 + rng.standard_normal((B, 21, 3)) * 0.01 + 0.02).astype(np.float32)
p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
rnge = np.arange(20, 51, 5)


def device_pass():
    a = EM.batch_compute_similarity_transform_torch(p, t)
    pck = EM.cal_PCK(a, t, rnge)                      # includes the 7-count device -> host read
    e = EM.mpjpe(a, t)
    acc = EM._accel(a, t)
    return pck, e, acc


for _ in range(3):
    device_pass()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    pck, e, acc = device_pass()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20

# the four kernels alone
a = EM.batch_compute_similarity_transform_torch(p, t)
parts = {}
for name, fn in (("procrustes", lambda: EM.batch_compute_similarity_transform_torch(p, t)),
                 ("joint_errors", lambda: EM._joint_errors(a, t, rnge, True)), ("accel", lambda: EM._accel(a, t))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    parts[name + "_us"] = e0.elapsed_time(e1) / 20 * 1e3

n_cpu = 8192
torch.set_num_threads(os.cpu_count() or 1)
pc, tc = torch.from_numpy(pred[:n_cpu]), torch.from_numpy(gt[:n_cpu])
w0 = time.perf_counter()
ac = eval_oracle.similarity_transform(pc, tc)
eval_oracle.cal_pck(ac, tc, rnge)
torch.sqrt(((ac - tc) ** 2).sum(-1)).mean(-1)
eval_oracle.compute_error_accel(tc.numpy(), ac.numpy())
cpu_s = time.perf_counter() - w0
print(json.dumps({"metric": "eval_metric_samples_per_s", "batch": B, "value": B / (ms * 1e-3), "ms_per_pass": ms, **parts,
                  "bytes_per_sample": 3 * 252 + 4, "hbm_gbs": B * (3 * 252 + 4) / (parts["procrustes_us"] * 1e-6) / 1e9,
                  "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{n_cpu} samples, oracle port of eval.py:110-161,300-316 + eval_utils.py:20-47"},
                  "pck_at_20mm": float(pck[0, -1])}))
