"""Evaluation metrics on the device (csrc/eval_metrics.cu, scat_b200/eval_metrics.py) against the fixture generated
from the reference's own functions (eval.py:110-161,300-340; data_utils/eval_utils.py:6-47) and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle
from scat_b200 import eval_metrics as EM

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_metrics.npz")


def test_metrics_match_reference_fixture():
    g = np.load(GOLD)
    pred, gt = torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["gt"]).cuda()
    aligned = EM.batch_compute_similarity_transform_torch(pred, gt)
    assert float((aligned.cpu() - torch.from_numpy(g["aligned"])).abs().max()) < 5e-7
    rnge = g["rnge"]
    for tag, p in (("raw", pred), ("pa", torch.from_numpy(g["aligned"]).cuda())):
        pck = EM.cal_PCK(p, gt, rnge)
        assert np.array_equal(pck, g[f"pck_{tag}"])                       # threshold counts: bit-exact
        assert abs(EM._area_under_curve(rnge / rnge.max(), pck[:, -1]) - float(g[f"auc_{tag}"])) < 1e-12
    assert np.allclose(EM.compute_accel(g["pred"]), g["accel"], rtol=2e-6, atol=0)
    assert np.allclose(EM.compute_error_accel(g["gt"], g["pred"]), g["accel_err"], rtol=2e-6, atol=0)
    assert np.allclose(EM.compute_error_accel(g["gt"], g["pred"], g["vis"]), g["accel_err_vis"], rtol=2e-6, atol=0)


def test_metrics_large_batch_against_oracle():
    rng = np.random.default_rng(11)
    B = 20000                                                             # ragged last block (20000 = 312 * 64 + 32)
    gt = (rng.standard_normal((B, 21, 3)) * 0.04).astype(np.float32)
    pred = (gt * 1.2 + rng.standard_normal((B, 21, 3)) * 0.015 + 0.03).astype(np.float32)
    p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    aligned = EM.batch_compute_similarity_transform_torch(p, t)
    ref = eval_oracle.similarity_transform(torch.from_numpy(pred), torch.from_numpy(gt))
    assert float((aligned.cpu() - ref).abs().max()) < 1e-6
    rnge = np.array([5.0, 12.5, 20, 35, 50])
    assert np.array_equal(EM.cal_PCK(p, t, rnge), eval_oracle.cal_pck(torch.from_numpy(pred), torch.from_numpy(gt), rnge))
    err = EM.mpjpe(aligned, t).cpu()
    want = torch.sqrt(((ref - torch.from_numpy(gt)) ** 2).sum(-1)).mean(-1)
    assert float((err - want).abs().max()) < 1e-6
    assert float(err.mean()) < float(EM.mpjpe(p, t).mean())              # alignment can only help
    acc = EM.compute_error_accel(gt[:500], pred[:500])
    assert np.allclose(acc, eval_oracle.compute_error_accel(gt[:500], pred[:500]), rtol=2e-6, atol=0)


def test_metrics_edge_cases_and_errors():
    one = torch.randn(1, 21, 3, device="cuda")
    out = EM.batch_compute_similarity_transform_torch(one * 2 + 1, one)
    assert float((out - one).abs().max()) < 1e-5                          # a pure similarity transform is undone exactly
    assert EM.compute_accel(torch.zeros(3, 21, 3, device="cuda")).shape == (1,)
    with pytest.raises(ValueError):
        EM.compute_accel(torch.zeros(2, 21, 3, device="cuda"))
    with pytest.raises(RuntimeError):
        EM.cal_PCK(torch.zeros(2, 21, 3), torch.zeros(2, 21, 3), [20])    # CPU tensors: no fallback
    with pytest.raises(ValueError):
        EM.batch_compute_similarity_transform_torch(one, torch.randn(1, 20, 3, device="cuda"))
