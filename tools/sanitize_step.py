#!/usr/bin/env python
"""Small programs for tools/sanitize.sh: a B = 8 train step (tf32 and bf16 precision, fp32 and bf16 seam, un-graphed so
that every launch is visible to the sanitizer), the coarse forward and the LBS forward + backward at B = 9."""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from scat_b200 import synth
from scat_b200.hand_net import EncoderTransformer, EncoderTransformerCoarse
from scat_b200.mano import ManoLayer
from scat_b200.train_step import HeadTrainStep

B = 8
mean = torch.from_numpy(synth.make_mean_params("hand"))
opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
x2, mf, labels = (torch.from_numpy(a).cuda() for a in synth.make_head_inputs(B, 0))
for precision, seam in (("tf32", "fp32"), ("tf32", "bf16"), ("bf16", "bf16")):
    net = EncoderTransformer(opt, mean, precision=precision, backbone=torch.nn.Identity())
    sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    ts = HeadTrainStep(net, B, use_graph=False, x2_dtype=seam)
    ts.load_inputs(x2.bfloat16() if seam == "bf16" else x2, mf, labels)
    random.seed(0)
    for _ in range(2):
        ts.set_mask()
        losses = ts.step()
    torch.cuda.synchronize()
    print(f"sanitize-step {precision}/{seam}: loss {float(losses[0]):.4f}", flush=True)
optc = SimpleNamespace(vit_heads=8, pl_reg=False, iteration=3, pos_embed=True, mask_rate=0.2)
netc = EncoderTransformerCoarse(optc, mean, precision="tf32", backbone=torch.nn.Identity())
sd = {k: torch.from_numpy(v) for k, v in synth.make_coarse_weights().items()}
sd["positionalEncoding.pe"] = netc.positionalEncoding.pe
netc.load_state_dict(sd, strict=True)
netc = netc.cuda()
with torch.no_grad():
    pred, fv, attn = netc.forward_features(mf, x2)
torch.cuda.synchronize()
print(f"sanitize-step coarse: pred sum {float(pred.sum()):.4f}", flush=True)
layer = ManoLayer(synth.make_mano_asset())
r, p, b = (torch.from_numpy(a).cuda().requires_grad_(True) for a in synth.make_mano_inputs(9, 1))
out = layer(r, p, b)
out.square().sum().backward()
torch.cuda.synchronize()
print(f"sanitize-step lbs: out sum {float(out.sum()):.4f}, grad sum {float(p.grad.sum()):.4f}", flush=True)
