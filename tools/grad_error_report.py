#!/usr/bin/env python
"""Per-tensor gradient error of the head at BASELINE config 2 (B=96) against the fp64 oracle, for each precision and
seam dtype: relative L2 and max-norm (max |err| / max |ref|).  Answers which tensors sit at the 1e-3 budget."""
import os, sys, random, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth
from tests.util import build_net, make_opt, oracle_step, rel_l2, rel_max
from tests.test_gpu_head import _run_module_step

out = {}
for precision, seam in (("tf32", "fp32"), ("tf32", "bf16"), ("bf16", "fp32"), ("bf16", "bf16")):
    regime = "unit"
    opt = make_opt(8, True, 3, True, 0.2)
    W = synth.make_head_weights(8, regime=regime)
    net = build_net(opt, W, "hand", precision=precision)
    x2, mf, labels = synth.make_head_inputs(96, 11)
    if seam == "bf16":
        x2 = torch.from_numpy(x2).bfloat16().float().numpy()
    r = _run_module_step(net, x2, mf, labels, mask_seed=3, seam=seam)
    o = oracle_step(W, x2, mf, labels, "hand", heads=8, iteration=3, pos_embed=True, mask_idx=net.last_mask, pl_reg=True,
                    dtype=torch.float64)
    named = dict(net.named_parameters())
    rows = {k: (rel_l2(named[k].grad, o["grads"][k]), rel_max(named[k].grad, o["grads"][k])) for k in W}
    rows["x2"] = (rel_l2(r["x2_grad"], o["x2_grad"]), rel_max(r["x2_grad"], o["x2_grad"]))
    rows["main_feat"] = (rel_l2(r["mf_grad"], o["main_feat_grad"]), rel_max(r["mf_grad"], o["main_feat_grad"]))
    rows["pred(out)"] = (rel_l2(r["pred"], o["pred"]), rel_max(r["pred"], o["pred"]))
    rows["feat_visual(out)"] = (rel_l2(r["fv"], o["feat_visual"]), rel_max(r["fv"], o["feat_visual"]))
    rows["pl(out)"] = (rel_l2(r["pl"], o["pl"]), rel_max(r["pl"], o["pl"]))
    print(f"=== precision {precision}, x2 seam {seam} ===")
    for k, (l2, mx) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:52s} rel-L2 {l2:.2e}   max-norm {mx:.2e}")
    out[f"{precision}/{seam}"] = {k: {"rel_l2": v[0], "max_norm": v[1]} for k, v in rows.items()}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_grad_error_report.json", "w"), indent=1)
