#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbs.py -x -q -m gpu > gpurun_out/r2lb_pytest_lbs.log 2>&1; echo "pytest lbs rc=$?" > gpurun_out/r2lb_rc.log
timeout 300 python - > gpurun_out/r2lb_bwd.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from scat_b200 import synth
from scat_b200.mano import ManoLayer
g = torch.Generator(device="cuda").manual_seed(1)
Bs = 16384
r = (0.5 * torch.randn(Bs, 3, device="cuda", generator=g)).requires_grad_(True)
p = (0.3 * torch.randn(Bs, 45, device="cuda", generator=g)).requires_grad_(True)
b = torch.randn(Bs, 10, device="cuda", generator=g).requires_grad_(True)
layer = ManoLayer(synth.make_mano_asset())
out = layer(r, p, b); go = torch.randn_like(out)
for _ in range(2): torch.autograd.grad(out, (r, p, b), go, retain_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): torch.autograd.grad(out, (r, p, b), go, retain_graph=True)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
print(f"LBS_BWD B={Bs}: {t:.3f} ms = {Bs / t / 1e3:.2f} M samples/s")
PY
timeout 300 python tools/lbs_quick.py > gpurun_out/r2lb_lbs_quick.log 2>&1
cat gpurun_out/r2lb_rc.log; tail -n 2 gpurun_out/r2lb_pytest_lbs.log; cat gpurun_out/r2lb_bwd.log | tail -2; grep -h LBS_QUICK gpurun_out/r2lb_lbs_quick.log
