"""Parameter containers mirroring the reference's models/vision_transformer_attn.py:14-113 module tree (the transformer of
``--net reg_transformer_coarse``, which also returns the last layer's attention maps).

Same contract as scat_b200/vision_transformer.py: reference class names, constructor arguments and attribute layout, so
``state_dict()`` keys and shapes are identical (``transformer.layers.{i}.0.to_qkv.weight``,
``transformer.layers.{i}.1.norm.weight`` for the post-attention LayerNorm, ``transformer.layers.{i}.2.fn.net.0.weight``,
``transformer.layers.2.2.net.0.weight`` for the last, norm-less feed-forward); parameters only, the arithmetic runs in the
sm_100a kernels behind ``scat_b200.hand_net.EncoderTransformerCoarse``.
"""
from __future__ import annotations

from torch import nn

from .vision_transformer import Attention, FeedForward, PreNorm, _no_eager


class PreNormAttn(nn.Module):        # vision_transformer_attn.py:21-26: a bare LayerNorm applied to the attention OUTPUT
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x, **kwargs):
        _no_eager("PreNormAttn")


class Transformer(nn.Module):        # vision_transformer_attn.py:88-113
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        if depth != 3:
            raise ValueError("scat_b200 builds the depth-3 narrowing transformer of hand_net.py:236")
        self.dim, self.depth, self.heads = dim, depth, heads
        self.layers = nn.ModuleList([])
        for i in range(depth):
            attn = Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)
            if i == depth - 1:       # :93-97
                self.layers.append(nn.ModuleList([attn, PreNormAttn(dim), FeedForward(dim, (dim * 3) // 4, out_dim=3)]))
            else:                    # :99-104 (mlp_dim is ignored: hidden = 3*dim//4)
                self.layers.append(nn.ModuleList([attn, PreNormAttn(dim), PreNorm(dim, FeedForward(dim, (dim * 3) // 4))]))
                dim = dim // 2

    def forward(self, x, mask=None):
        _no_eager("Transformer (attention-map variant)")

    def ordered_parameters(self):
        """The 31 transformer tensors in the C ABI's head order (include/scat_b200.h): the slots norm_a.w / norm_a.b carry
        the post-attention LayerNorm (layers.i.1.norm)."""
        out = []
        for i, (attn, pren, ff) in enumerate(self.layers):
            out += [pren.norm.weight, pren.norm.bias, attn.to_qkv.weight, attn.to_out[0].weight, attn.to_out[0].bias]
            if i < self.depth - 1:
                out += [ff.norm.weight, ff.norm.bias, ff.fn.net[0].weight, ff.fn.net[0].bias, ff.fn.net[2].weight,
                        ff.fn.net[2].bias]
            else:
                out += [ff.net[0].weight, ff.net[0].bias, ff.net[2].weight, ff.net[2].bias]
        return out
