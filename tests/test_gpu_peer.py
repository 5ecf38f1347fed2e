"""NVLink peer-memory gradient all-reduce (csrc/dp_allreduce.cu, dp.PeerMemory).  The single-GPU cases run
everywhere; the two-rank comparison against NCCL needs two GPUs and is skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_allreduce_is_identity_and_advances_epoch():
    from scat_b200 import dp
    from scat_b200._lib import stream_ptr
    pm = dp.PeerMemory(10_001, "cuda:0")
    assert pm.world == 1 and pm.n_pad == 10_004
    x = torch.randn(10_001, device="cuda")
    pm.flat.copy_(x)
    for _ in range(3):                                   # barriers with itself: epochs 0, 1, 2
        pm.enqueue(stream_ptr())
    pm.enqueue(stream_ptr(), 8, 4000)
    torch.cuda.synchronize()
    assert not pm.timed_out()
    assert torch.equal(pm.flat, x)
    g = torch.cuda.CUDAGraph()                           # capturable
    with torch.cuda.graph(g):
        pm.enqueue(stream_ptr())
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(pm.flat, x) and not pm.timed_out()
    pm.close()


def test_allreduce_rejects_bad_ranges():
    from scat_b200 import dp
    from scat_b200._lib import stream_ptr
    pm = dp.PeerMemory(64, "cuda:0")
    for lo, hi in ((2, 8), (0, 6), (8, 8), (-4, 8)):
        with pytest.raises(RuntimeError):
            pm.enqueue(stream_ptr(), lo, hi)
    pm.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_two_ranks_match_nccl_and_train_step():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
