"""Shared helpers for the parity tests."""
import os
import random
from types import SimpleNamespace

import numpy as np
import torch

from scat_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_max(a, b):
    """max |a-b| / max |b|"""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


class StubBackbone(torch.nn.Module):
    """Stands in for the ResNet-50 feature producer: returns fixed seam tensors (resnet.py:162 signature)."""

    def __init__(self):
        super().__init__()
        self.main_feat = None
        self.x2 = None

    def forward(self, _img):
        return self.main_feat, None, self.x2, None, None


def make_opt(heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2):
    return SimpleNamespace(vit_heads=heads, pl_reg=pl_reg, iteration=iteration, pos_embed=pos_embed,
                           mask_rate=mask_rate)


def build_net(opt, weights, mean_kind="hand", precision="fp32", device="cuda"):
    """scat_b200 EncoderTransformer with a stub backbone and the given synthetic weights loaded strictly
    (apart from the backbone) through load_state_dict, exactly as a reference checkpoint would be."""
    from scat_b200.hand_net import EncoderTransformer
    mean = torch.from_numpy(synth.make_mean_params(mean_kind))
    net = EncoderTransformer(opt, mean, precision=precision, backbone=StubBackbone())
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    sd["positionalEncoding.pe"] = net.positionalEncoding.pe
    missing, unexpected = net.load_state_dict(sd, strict=True), None
    return net.to(device)


def oracle_step(weights, x2, mf, labels, mean_kind, *, heads, iteration, pos_embed, mask_idx, pl_reg, dtype=torch.float32):
    from oracle import head_oracle
    P = {k: torch.from_numpy(v).to(dtype) for k, v in weights.items()}
    mean = torch.from_numpy(synth.make_mean_params(mean_kind)).to(dtype)
    return head_oracle.train_step(P, torch.from_numpy(x2).to(dtype), torch.from_numpy(mf).to(dtype),
                                  torch.from_numpy(labels).to(dtype), mean, heads=heads, iteration=iteration,
                                  pos_embed=pos_embed, mask_idx=mask_idx, pl_reg=pl_reg)
