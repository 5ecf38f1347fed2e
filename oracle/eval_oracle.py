"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's evaluation metrics, the step
after the head at inference.

  * similarity_transform   batch_compute_similarity_transform_torch   /root/reference/eval.py:110-161
  * cal_pck                cal_PCK                                     /root/reference/eval.py:300-316
  * area_under_curve       _area_under_curve                           /root/reference/eval.py:328-340
  * compute_accel          compute_accel                               /root/reference/data_utils/eval_utils.py:6-17
  * compute_error_accel    compute_error_accel                         /root/reference/data_utils/eval_utils.py:20-47

Parity pin: eval.py cannot be imported as a module (it imports files that are not in the repository, SURVEY.md
section 8c), so oracle/make_golden.py extracts exactly these function definitions from the reference's source files
with ``ast`` in the authoring container, executes them on seeded inputs, asserts agreement with this restatement and
stores tests/golden/eval_metrics.npz.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Reference quirks kept: cal_PCK fills every per-joint column with the SAME number (it takes ``dist.flat``, i.e. all
joints, inside the per-joint loop), so the "average" column equals it; distances are compared in millimetres
(inputs in metres, x1000); the AUC normalises by the trapezoid of ones over the (deduplicated) thresholds;
compute_error_accel drops a frame triple when any of its three frames is invisible.
"""
from __future__ import annotations

import numpy as np
import torch


def similarity_transform(S1: torch.Tensor, S2: torch.Tensor) -> torch.Tensor:
    """[B,J,3] predictions aligned onto [B,J,3] targets (the layout every call site uses, eval.py:953)."""
    X1t, X2t = S1.permute(0, 2, 1), S2.permute(0, 2, 1)               # [B,3,J]
    mu1, mu2 = X1t.mean(dim=-1, keepdim=True), X2t.mean(dim=-1, keepdim=True)
    X1, X2 = X1t - mu1, X2t - mu2
    var1 = (X1 ** 2).sum(dim=1).sum(dim=1)
    K = X1.bmm(X2.permute(0, 2, 1))
    U, s, Vh = torch.linalg.svd(K)                                    # torch.svd returns V, linalg.svd V^H
    V = Vh.transpose(1, 2)
    Z = torch.eye(3, dtype=S1.dtype).repeat(S1.shape[0], 1, 1)
    Z[:, -1, -1] *= torch.sign(torch.det(U.bmm(V.permute(0, 2, 1))))
    R = V.bmm(Z.bmm(U.permute(0, 2, 1)))
    scale = torch.einsum("bii->b", R.bmm(K)) / var1
    t = mu2 - scale[:, None, None] * R.bmm(mu1)
    return (scale[:, None, None] * R.bmm(X1t) + t).permute(0, 2, 1)


def cal_pck(pred: torch.Tensor, gt: torch.Tensor, rnge) -> np.ndarray:
    dist = torch.sqrt(((pred * 1000 - gt * 1000) ** 2).sum(dim=-1)).numpy()
    pck = np.zeros((len(rnge), dist.shape[1] + 1))
    for k, r in enumerate(rnge):
        pck[k, :-1] = 100.0 * np.mean(dist.flat <= r)
        pck[k, -1] = np.mean(pck[k, :-1])
    return pck


def area_under_curve(xpts: np.ndarray, ypts: np.ndarray) -> float:
    _, idx = np.unique(xpts, return_index=True)
    x, y = xpts[idx], ypts[idx]
    trap = getattr(np, "trapezoid", None) or np.trapz
    return float(trap(y, x) / trap(np.ones_like(x), x))


def compute_accel(joints: np.ndarray) -> np.ndarray:
    vel = joints[1:] - joints[:-1]
    acc = vel[1:] - vel[:-1]
    return np.mean(np.linalg.norm(acc, axis=2), axis=1)


def compute_error_accel(joints_gt: np.ndarray, joints_pred: np.ndarray, vis=None) -> np.ndarray:
    a_gt = joints_gt[:-2] - 2 * joints_gt[1:-1] + joints_gt[2:]
    a_pr = joints_pred[:-2] - 2 * joints_pred[1:-1] + joints_pred[2:]
    normed = np.linalg.norm(a_pr - a_gt, axis=2)
    if vis is None:
        keep = np.ones(len(normed), dtype=bool)
    else:
        invis = np.logical_not(vis)
        bad = np.logical_or(invis, np.logical_or(np.roll(invis, -1), np.roll(invis, -2)))[:-2]
        keep = np.logical_not(bad)
    return np.mean(normed[keep], axis=1)
