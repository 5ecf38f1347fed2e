"""Parameter containers mirroring the reference's models/vision_transformer.py:13-101 module tree.

The classes keep the reference's names, constructor arguments and attribute layout so that
``state_dict()`` keys and shapes are identical (``transformer.layers.{i}.0.fn.norm.weight`` ...,
``transformer.layers.2.1.net.0.weight`` for the last, norm-less feed-forward) and existing ``.pth`` files load
with ``strict=True``.  They hold parameters only: the arithmetic runs in the sm_100a kernels behind
``scat_b200.hand_net.EncoderTransformer`` / ``scat_b200.functional``; calling ``forward`` on a container is
an error rather than a silent PyTorch fallback.
"""
from __future__ import annotations

import torch
from torch import nn


def _no_eager(name):
    raise RuntimeError(
        f"scat_b200.vision_transformer.{name} holds parameters only; run the model through "
        f"scat_b200.hand_net.EncoderTransformer or scat_b200.functional (CUDA kernels, no PyTorch fallback)")


class Residual(nn.Module):           # vision_transformer.py:13-18
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        _no_eager("Residual")


class PreNorm(nn.Module):            # vision_transformer.py:20-26
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        _no_eager("PreNorm")


class FeedForward(nn.Module):        # vision_transformer.py:28-44
    def __init__(self, dim, hidden_dim, out_dim=None):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Linear(hidden_dim, dim // 2 if out_dim is None else 3),
        )

    def forward(self, x):
        _no_eager("FeedForward")


class Attention(nn.Module):          # vision_transformer.py:46-79
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        if dim_head != 64:
            raise ValueError("scat_b200 attention kernels are built for dim_head=64 (hand_net.py:331)")
        if dropout != 0.0:
            raise ValueError("dropout is 0.0 everywhere on the reference path (hand_net.py:331)")
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def forward(self, x, mask=None):
        _no_eager("Attention")


class Transformer(nn.Module):        # vision_transformer.py:81-101
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        if depth != 3:
            raise ValueError("scat_b200 builds the depth-3 narrowing transformer of hand_net.py:331")
        self.dim, self.depth, self.heads = dim, depth, heads
        self.layers = nn.ModuleList([])
        for i in range(depth):
            attn = Residual(PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)))
            if i == depth - 1:       # last layer: bare FeedForward to 3 outputs, no PreNorm (:86-90)
                self.layers.append(nn.ModuleList([attn, FeedForward(dim, (dim * 3) // 4, out_dim=3)]))
            else:                    # mlp_dim is ignored by the reference: hidden = 3*dim//4 (:94)
                self.layers.append(nn.ModuleList([attn, PreNorm(dim, FeedForward(dim, (dim * 3) // 4))]))
                dim = dim // 2

    def forward(self, x, mask=None):
        """x[B,n,dim] -> [B,n,3] through the CUDA kernels (mask must be None like hand_net.py:375)."""
        if mask is not None:
            raise NotImplementedError("the reference path always calls the transformer with mask=None")
        from . import functional
        return functional.token_transformer(self, x)

    def ordered_parameters(self):
        """The 31 transformer tensors in the C ABI order (include/scat_b200.h)."""
        out = []
        for i, (attn, ff) in enumerate(self.layers):
            pre = attn.fn
            out += [pre.norm.weight, pre.norm.bias, pre.fn.to_qkv.weight, pre.fn.to_out[0].weight, pre.fn.to_out[0].bias]
            if i < self.depth - 1:
                out += [ff.norm.weight, ff.norm.bias, ff.fn.net[0].weight, ff.fn.net[0].bias, ff.fn.net[2].weight,
                        ff.fn.net[2].bias]
            else:
                out += [ff.net[0].weight, ff.net[0].bias, ff.net[2].weight, ff.net[2].bias]
        return out
