"""Evaluation metrics of the reference's eval loop on the device, under the reference's
own function names so ``eval.py`` can switch imports:

    batch_compute_similarity_transform_torch(S1, S2)    eval.py:110-161   Procrustes alignment (3x3 SVD per sample)
    cal_PCK(pred_joints, gt_joints, rnge)               eval.py:300-316   PCK table [len(rnge), J + 1]
    _area_under_curve(xpts, ypts)                       eval.py:328-340   AUC of the PCK curve (7 numbers: host)
    compute_accel(joints)                               data_utils/eval_utils.py:6-17
    compute_error_accel(joints_gt, joints_pred, vis)    data_utils/eval_utils.py:20-47
    mpjpe(pred_joints, gt_joints)                       eval.py:749       per-sample mean joint error

Inputs are CUDA tensors [B, J, 3] (numpy arrays are moved to the current device); the reference pulls every frame to
numpy instead (eval.py:691-753).  There is no CPU path: CPU tensors raise.  Reference quirks kept: cal_PCK fills
every per-joint column with the all-joint value (eval.py:309-312 uses ``dist.flat`` inside the per-joint loop) and
compares millimetres (inputs are metres).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _joints(x, name: str) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
        raise RuntimeError(f"{name}: CUDA tensor (or numpy array) expected -- the metrics have no CPU path")
    if x.dim() != 3 or x.shape[-1] != 3:
        raise ValueError(f"{name}: expected [B, J, 3], got {tuple(x.shape)}")
    return x.detach().to(torch.float32).contiguous()


def batch_compute_similarity_transform_torch(S1, S2) -> torch.Tensor:
    """``S1`` [B,J,3] aligned onto ``S2`` by the optimal similarity transform, per sample."""
    s1, s2 = _joints(S1, "S1"), _joints(S2, "S2")
    if s1.shape != s2.shape:
        raise ValueError(f"S1 {tuple(s1.shape)} and S2 {tuple(s2.shape)} differ")
    out = torch.empty_like(s1)
    check(_lib.load().scat_eval_procrustes(ptr(s1), ptr(s2), s1.shape[0], s1.shape[1], ptr(out), None, stream_ptr()),
          "scat_eval_procrustes")
    return out


def _joint_errors(pred, gt, rnge, want_mpjpe: bool):
    p, g = _joints(pred, "pred_joints"), _joints(gt, "gt_joints")
    if p.shape != g.shape:
        raise ValueError(f"pred {tuple(p.shape)} and gt {tuple(g.shape)} differ")
    thr = np.asarray(rnge, dtype=np.float64).reshape(-1)
    counts = torch.empty(max(len(thr), 1), dtype=torch.int64, device=p.device)
    err = torch.empty(p.shape[0], device=p.device) if want_mpjpe else None
    arr = (C.c_double * max(len(thr), 1))(*thr.tolist())
    check(_lib.load().scat_eval_joint_errors(ptr(p), ptr(g), p.shape[0], p.shape[1], 1000.0, arr, len(thr), ptr(counts),
                                             ptr(err), stream_ptr()), "scat_eval_joint_errors")
    return counts[: len(thr)], err, p.shape[0] * p.shape[1], p.shape[1]


def cal_PCK(pred_joints, gt_joints, rnge) -> np.ndarray:
    counts, _, total, n_joints = _joint_errors(pred_joints, gt_joints, rnge, False)
    frac = 100.0 * (counts.cpu().numpy().astype(np.float64) / total)          # 100 * np.mean(dist.flat <= rngval)
    pck = np.repeat(frac[:, None], n_joints + 1, axis=1)
    pck[:, -1] = np.mean(pck[:, :-1], axis=1)
    return pck


def mpjpe(pred_joints, gt_joints) -> torch.Tensor:
    """Per-sample mean joint distance (device tensor [B]), ``torch.sqrt(((p - g) ** 2).sum(-1)).mean(-1)``."""
    return _joint_errors(pred_joints, gt_joints, [], True)[1]


def _area_under_curve(xpts, ypts) -> float:
    xpts, ypts = np.asarray(xpts, dtype=np.float64), np.asarray(ypts, dtype=np.float64)
    _, idx = np.unique(xpts, return_index=True)
    x, y = xpts[idx], ypts[idx]
    trap = getattr(np, "trapezoid", None) or np.trapz
    return float(trap(y, x) / trap(np.ones_like(x), x))


def _accel(pred, gt):
    p = _joints(pred, "joints_pred")
    g = None if gt is None else _joints(gt, "joints_gt")
    if g is not None and g.shape != p.shape:
        raise ValueError(f"pred {tuple(p.shape)} and gt {tuple(g.shape)} differ")
    if p.shape[0] < 3:
        raise ValueError("acceleration needs at least 3 frames")
    out = torch.empty(p.shape[0] - 2, device=p.device)
    check(_lib.load().scat_eval_accel(ptr(p), ptr(g), p.shape[0], p.shape[1], ptr(out), stream_ptr()), "scat_eval_accel")
    return out


def compute_accel(joints) -> np.ndarray:
    return _accel(joints, None).cpu().numpy()


def compute_error_accel(joints_gt, joints_pred, vis=None) -> np.ndarray:
    err = _accel(joints_pred, joints_gt).cpu().numpy()
    if vis is None:
        return err
    invis = np.logical_not(np.asarray(vis, dtype=bool))
    bad = np.logical_or(invis, np.logical_or(np.roll(invis, -1), np.roll(invis, -2)))[:-2]
    return err[np.logical_not(bad)]
