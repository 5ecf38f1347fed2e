"""Fused Adam for the head (csrc/adam.cu, scat_b200/optim.py) against torch.optim.Adam -- the reference's optimiser
(train.py:60,209) -- through the committed fixture, on the device, inside the step's CUDA graph, and through
state_dict round trips."""
import os

import numpy as np
import pytest
import torch

from scat_b200 import dp
from tests.test_gpu_head import _config2
from tests.util import rel_max

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag,wd", [("wd0", 0.0), ("wd1", 0.01)])
def test_adam_kernel_matches_torch_adam_fixture(tag, wd):
    from scat_b200.optim import HeadAdam
    g = np.load(os.path.join(GOLD, "adam.npz"))
    n = int(g["n_tensors"])
    ps = [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).cuda()) for i in range(n)]
    opt = HeadAdam(ps, lr=float(g["lrs"][0]), weight_decay=wd)
    bucket = dp.FlatGradBucket(ps)
    for k in range(6):
        opt.param_groups[0]["lr"] = float(g["lrs"][k])
        for i in range(n):
            bucket.views[i].copy_(torch.from_numpy(g[f"g{k}_{i}"]))
        opt.step()
    torch.cuda.synchronize()
    for i, p in enumerate(ps):
        st = opt.state[p]
        for got, key in ((p.data, "p"), (st["exp_avg"], "m"), (st["exp_avg_sq"], "v")):
            assert rel_max(got, g[f"{tag}_{key}{i}"]) <= 2e-7, (key, i)


def test_head_adam_equals_torch_adam_on_head_and_state_dict_round_trip():
    from scat_b200.optim import HeadAdam
    from scat_b200.train_step import HeadTrainStep
    opt_cfg, W, net, x2, mf, labels = _config2("tf32", B=8)
    ts = HeadTrainStep(net, 8)
    ts.load_inputs(torch.from_numpy(x2).cuda(), torch.from_numpy(mf).cuda(), torch.from_numpy(labels).cuda())
    ts.set_mask(list(range(ts.n_masked)))
    fused = HeadAdam(net.head_parameters(), lr=1e-4)
    twins = [torch.nn.Parameter(p.detach().clone()) for p in net.head_parameters()]
    ref = torch.optim.Adam(twins, lr=1e-4)
    for k in range(4):
        lr = 1e-4 * (k + 1) / 15
        fused.param_groups[0]["lr"] = lr
        ref.param_groups[0]["lr"] = lr
        ts.step()
        for t, p in zip(twins, net.head_parameters()):
            t.grad = p.grad.detach().clone()
        fused.step()
        ref.step()
        if k == 1:                       # checkpoint through torch.optim.Adam's own layout and back
            sd = fused.state_dict()
            probe = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in net.head_parameters()], lr=1.0)
            probe.load_state_dict(sd)    # torch accepts it
            fused.load_state_dict(probe.state_dict())
            assert fused.step_count == 2
    torch.cuda.synchronize()
    for t, p in zip(twins, net.head_parameters()):
        assert rel_max(p.data, t.data) < 1e-6
        assert rel_max(fused.state[p]["exp_avg"], ref.state[t]["exp_avg"]) < 1e-6
        assert rel_max(fused.state[p]["exp_avg_sq"], ref.state[t]["exp_avg_sq"]) < 1e-6


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_with_optimizer_in_graph(use_graph):
    """step(optimize=True): forward, backward and the Adam update in one CUDA graph, the learning rate following the
    reference's warm-up schedule from device memory.  Checked against torch.optim.Adam fed the same gradients."""
    from scat_b200.optim import HeadAdam, WarmupSchedule
    from scat_b200.train_step import HeadTrainStep
    opt_cfg, W, net, x2, mf, labels = _config2("tf32", B=8)
    ts = HeadTrainStep(net, 8, use_graph=use_graph)
    ts.load_inputs(torch.from_numpy(x2).cuda(), torch.from_numpy(mf).cuda(), torch.from_numpy(labels).cuda())
    fused = HeadAdam(net.head_parameters(), lr=1e-4)
    sched = WarmupSchedule(fused)
    ts.attach_optimizer(fused)
    twins = [torch.nn.Parameter(p.detach().clone()) for p in net.head_parameters()]
    ref = torch.optim.Adam(twins, lr=1e-4)
    before = ts.flat_params.clone()
    losses = []
    for epoch in range(3):
        sched.step(epoch + 1)
        ref.param_groups[0]["lr"] = fused.param_groups[0]["lr"]
        for _ in range(2):
            ts.set_mask(list(range(ts.n_masked)))
            losses.append(float(ts.step(optimize=True)[0]))
            for t, v in zip(twins, ts.bucket.views):
                t.grad = v.detach().clone()
            ref.step()
    torch.cuda.synchronize()
    assert fused.step_count == 6
    assert float((ts.flat_params - before).abs().max()) > 1e-5           # the weights moved ...
    assert losses[-1] < losses[0]                                         # ... downhill
    for t, p in zip(twins, net.head_parameters()):
        assert rel_max(p.data, t.data) < 1e-6
