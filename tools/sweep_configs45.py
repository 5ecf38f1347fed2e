#!/usr/bin/env python
"""BASELINE configs 4 and 5 measured on the GPU (CUDA events, 20 iterations after 3 warm-ups):
  config 5: MANO LBS forward and the autoregressive regressor (both head shapes), batch 1k..64k, as samples/s and as
            achieved GB/s of algorithmic bytes (LBS 9,820 B/sample, regressor 4,624 B/sample)
            against the measured HBM peak;
  config 4: the n=128 x dim=196 token transformer, batch 256, inference (fp32 / tf32 / bf16).
Prints one JSON object; profiles/r1_configs45.json is a committed run."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import synth, functional as SF
from scat_b200.mano import ManoLayer

HBM = 6454.6
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    HBM = float(json.load(open(p))["hbm_gbs"])


def t_us(fn, it=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


res = {"hbm_peak_gbs": HBM, "lbs": [], "regressor_head": [], "regressor_h3dw": [], "config4": []}
layer = ManoLayer(synth.make_mano_asset())
W = synth.make_head_weights(8)
wr, br = torch.from_numpy(W["regressor.weight"]).cuda(), torch.from_numpy(W["regressor.bias"]).cuda()
mean = torch.from_numpy(synth.make_mean_params("hand")).cuda()
g = torch.Generator(device="cuda").manual_seed(1)
wr61 = 0.02 * torch.randn(61, 1085, device="cuda", generator=g); br61 = torch.zeros(61, device="cuda")
for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    rots = 0.5 * torch.randn(B, 3, device="cuda", generator=g); poses = 0.3 * torch.randn(B, 45, device="cuda", generator=g)
    betas = torch.randn(B, 10, device="cuda", generator=g)
    out = torch.empty(B, 799, 3, device="cuda")
    us = t_us(lambda: layer(rots, poses, betas, out=out))
    res["lbs"].append({"batch": B, "us": us, "samples_per_s": B / us * 1e6, "gbs": 9820 * B / us / 1e3, "frac_hbm": 9820 * B / us / 1e3 / HBM})
    mf = torch.relu(torch.randn(B, 1024, device="cuda", generator=g)); fo = 0.05 * torch.randn(B, 63, device="cuda", generator=g)
    us = t_us(lambda: SF.regressor_fwd(mf, fo, mean, wr, br, iteration=3, root_relative=True))
    res["regressor_head"].append({"batch": B, "us": us, "samples_per_s": B / us * 1e6, "gbs": 4624 * B / us / 1e3, "frac_hbm": 4624 * B / us / 1e3 / HBM})
    us = t_us(lambda: SF.regressor_fwd(mf, None, torch.zeros(1, 61, device="cuda"), wr61, br61, iteration=3, root_relative=False))
    bytes61 = 4096 + 2 * 244
    res["regressor_h3dw"].append({"batch": B, "us": us, "samples_per_s": B / us * 1e6, "gbs": bytes61 * B / us / 1e3, "frac_hbm": bytes61 * B / us / 1e3 / HBM})

from scat_b200.vision_transformer import Transformer
from scat_b200.hand_net import PositionalEncoding
B, n, dim, heads = 256, 128, 196, 8
Wt = synth.make_token_weights(dim, heads)
tr = Transformer(dim=dim, depth=3, heads=heads, dim_head=64, mlp_dim=2 * dim)
tr.load_state_dict({k[len("transformer."):]: torch.from_numpy(v) for k, v in Wt.items() if k.startswith("transformer.")})
tr = tr.cuda()
pe = PositionalEncoding(dim, max_len=n).pe[0].cuda()
tok = torch.from_numpy(synth.make_token_inputs(B, n, dim, 5)).cuda()
idx = torch.tensor(list(range(0, 25)), dtype=torch.int32, device="cuda")
mt = torch.from_numpy(Wt["mask_token"]).cuda().view(-1)
with torch.no_grad():
    for prec in ("fp32", "tf32", "bf16"):
        us = t_us(lambda: SF.token_transformer(tr, tok, mask_token=mt, pe=pe, mask_idx=idx, precision=prec), it=10)
        res["config4"].append({"precision": prec, "batch": B, "us": us, "samples_per_s": B / us * 1e6,
                               "tflops": 297.4e6 * B / us / 1e6})
print(json.dumps(res, indent=1))
