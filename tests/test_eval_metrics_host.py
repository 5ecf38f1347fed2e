"""CPU checks of the evaluation-metric row: the oracle against the fixture made from the
reference's own functions, and the SHARED device/host arithmetic (scat_b200/csrc/eval_math.cuh, compiled here with
g++) against both."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from oracle import eval_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "eval_metrics.npz")


def test_eval_oracle_matches_reference_fixture():
    g = np.load(GOLD)
    pred, gt = torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"])
    aligned = eval_oracle.similarity_transform(pred, gt)
    assert float((aligned - torch.from_numpy(g["aligned"])).abs().max()) < 2e-6
    rnge = g["rnge"]
    for tag, p in (("raw", pred), ("pa", torch.from_numpy(g["aligned"]))):
        pck = eval_oracle.cal_pck(p, gt, rnge)
        assert np.array_equal(pck, g[f"pck_{tag}"])
        assert abs(eval_oracle.area_under_curve(rnge / rnge.max(), pck[:, -1]) - float(g[f"auc_{tag}"])) < 1e-12
    assert np.array_equal(eval_oracle.compute_accel(g["pred"]), g["accel"])
    assert np.array_equal(eval_oracle.compute_error_accel(g["gt"], g["pred"]), g["accel_err"])
    assert np.array_equal(eval_oracle.compute_error_accel(g["gt"], g["pred"], g["vis"]), g["accel_err_vis"])
    assert g["accel_err_vis"].shape[0] == g["accel_err"].shape[0] - 6       # two invisible frames drop three triples each
    # alignment properties: the Procrustes-aligned error is never larger, and aligning a pure similarity transform is exact
    assert float(g["pck_pa"][0, -1]) >= float(g["pck_raw"][0, -1])


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("evalhost") / "eval_math_host")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "eval_math_host.cpp")], check=True)
    return exe


def _run_host(exe, pred, gt):
    B, n = pred.shape[:2]
    blob = struct.pack("ii", B, n) + pred.astype(np.float32).tobytes() + gt.astype(np.float32).tobytes()
    out = subprocess.run([exe], input=blob, capture_output=True, check=True).stdout
    aligned = np.frombuffer(out[: B * n * 12], dtype=np.float32).reshape(B, n, 3)
    return aligned, np.frombuffer(out[B * n * 12:], dtype=np.float32)


def test_shared_procrustes_arithmetic_on_host(host_binary):
    """eval_math.cuh::similarity_align (the function the CUDA kernel runs per thread) == the reference's result."""
    g = np.load(GOLD)
    aligned, scale = _run_host(host_binary, g["pred"], g["gt"])
    assert np.abs(aligned - g["aligned"]).max() < 5e-7
    assert np.all((scale > 0.5) & (scale < 1.6))
    rng = np.random.default_rng(3)
    p = (rng.standard_normal((3000, 21, 3)) * 0.05).astype(np.float32)
    t = (rng.standard_normal((3000, 21, 3)) * 0.05).astype(np.float32)
    a, _ = _run_host(host_binary, p, t)
    o = eval_oracle.similarity_transform(torch.from_numpy(p), torch.from_numpy(t)).numpy()
    assert np.abs(a - o).max() < 1e-6
    # exact recovery of a similarity transform, including one that needs the reflection fix and coplanar points
    t = rng.standard_normal((8, 21, 3)).astype(np.float32)
    t[4:, :, 2] = 0.0                                                   # coplanar targets: third singular value 0
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    p = (1.7 * t @ q.T + np.array([0.3, -0.2, 0.1])).astype(np.float32)
    a, s = _run_host(host_binary, p, t)
    assert np.abs(a - t).max() < 5e-6 and np.allclose(s, 1 / 1.7, rtol=1e-5)
    # different joint counts
    for n in (3, 5, 64):
        p = rng.standard_normal((10, n, 3)).astype(np.float32)
        t = rng.standard_normal((10, n, 3)).astype(np.float32)
        a, _ = _run_host(host_binary, p, t)
        o = eval_oracle.similarity_transform(torch.from_numpy(p), torch.from_numpy(t)).numpy()
        assert np.abs(a - o).max() < 2e-5, n
