"""ORACLE (test infrastructure, not product code): CPU restatement of SCAT's reg_transformer head.

This file restates, in plain functional PyTorch on the CPU, what the reference computes on the
hot path named by BASELINE.json:north_star:

  * EncoderTransformer.forward          /root/reference/models/hand_net.py:355-398
  * PositionalEncoding                  /root/reference/models/hand_net.py:61-77
  * Transformer / Attention / FF / LN   /root/reference/models/vision_transformer.py:13-101
  * projection + losses (train step)    /root/reference/train.py:112-120,165-203
  * H3DWEncoder regressor (adjacent)    /root/reference/models/hand_net.py:49-58

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it, and only as the checker / the timed CPU baseline.  The product (scat_b200/) never imports it.

Parity pin: oracle/make_golden.py imports the UNMODIFIED reference from /root/reference in the
authoring container, runs both on identical inputs, asserts agreement and writes tests/golden/*.npz;
tests/test_oracle_golden.py re-checks this restatement against those fixtures wherever it runs.

Everything is dtype-generic: feed float64 tensors to get a float64 "truth" for error budgeting of the
TF32 / BF16 kernels, float32 to reproduce the reference's own arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

DIM_HEAD = 64


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def positional_encoding(max_len: int, d_model: int, dtype=torch.float32) -> torch.Tensor:
    """pe[1,max_len,d_model]; hand_net.py:66-73 (computed in fp32 like the reference buffer)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).to(dtype)


def _layer_keys(i: int, last: bool):
    p = f"transformer.layers.{i}."
    ff = p + ("1.net." if last else "1.fn.net.")
    return p, ff


def attention(x, w_qkv, w_out, b_out, heads):
    """vision_transformer.py:59-79 with mask=None."""
    b, n, _ = x.shape
    qkv = F.linear(x, w_qkv)                                   # :61, no bias
    q, k, v = qkv.chunk(3, dim=-1)
    def split(t):                                              # 'b n (h d) -> b h n d', :62
        return t.reshape(b, n, heads, DIM_HEAD).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    dots = torch.matmul(q, k.transpose(-1, -2)) * (DIM_HEAD ** -0.5)   # :64
    attn = dots.softmax(dim=-1)                                # :74
    out = torch.matmul(attn, v)                                # :76
    out = out.permute(0, 2, 1, 3).reshape(b, n, heads * DIM_HEAD)      # :77
    return F.linear(out, w_out, b_out)                         # :78


def transformer(x, P: Dict[str, torch.Tensor], heads: int, depth: int = 3):
    """vision_transformer.py:97-101: x = attn(x) + x ; x = ff(norm(x))  (no residual around ff;
    the last layer's ff has no PreNorm and maps to 3, :86-90)."""
    for i in range(depth):
        last = i == depth - 1
        p, ff = _layer_keys(i, last)
        d = x.shape[-1]
        h = F.layer_norm(x, (d,), P[p + "0.fn.norm.weight"], P[p + "0.fn.norm.bias"], 1e-5)
        x = attention(h, P[p + "0.fn.fn.to_qkv.weight"], P[p + "0.fn.fn.to_out.0.weight"],
                      P[p + "0.fn.fn.to_out.0.bias"], heads) + x
        y = x if last else F.layer_norm(x, (d,), P[p + "1.norm.weight"], P[p + "1.norm.bias"], 1e-5)
        y = F.gelu(F.linear(y, P[ff + "0.weight"], P[ff + "0.bias"]))      # exact erf GELU, :33-34
        x = F.linear(y, P[ff + "2.weight"], P[ff + "2.bias"])
    return x


# --------------------------------------------------------------------------------------------
# the head
# --------------------------------------------------------------------------------------------
def head_forward(P: Dict[str, torch.Tensor], x2: torch.Tensor, main_feat: torch.Tensor,
                 mean_params: torch.Tensor, *, heads: int = 8, iteration: int = 3,
                 pos_embed: bool = True, mask_idx: Optional[Sequence[int]] = None,
                 pl_reg: bool = False, pe: Optional[torch.Tensor] = None):
    """hand_net.py:363-398 with the backbone outputs (main_feat, x2) as inputs.

    mask_idx is the host-drawn list of masked token indices (hand_net.py:370-372); [] / None = no
    masking.  Returns (pred_params[B,66], feat_visual[B,21,28,28][, pl_term[B,21,28,28]]).
    Reference quirks kept on purpose:
      * masking is applied regardless of train/eval mode (:369);
      * with pos_embed=False the masked overwrite is IN PLACE on a view of feat_visual (:364,:373),
        so the returned feat_visual carries mask_token rows and pl_term is taken w.r.t. that
        overwritten tensor;
      * joint 1 of the output is exactly zero (:389-393);
      * pl_term is a constant (no create_graph, :396).
    """
    B = x2.shape[0]
    n_tok = P["conv1x1_channel_reduction.weight"].shape[0]
    feat_visual = F.conv2d(x2, P["conv1x1_channel_reduction.weight"])          # :363
    feat = feat_visual.view(B, n_tok, -1)                                       # :364 (a view)
    if pos_embed:
        if pe is None:
            pe = positional_encoding(n_tok, feat.shape[-1], dtype=feat.dtype).to(feat.device)
        feat = feat + pe[: feat.size(0), :]                                     # :75-77 (slices dim 0 of size 1)
    if mask_idx is not None and len(mask_idx) > 0:
        idx = mask_idx if torch.is_tensor(mask_idx) else list(mask_idx)         # (an index tensor keeps a GPU run sync-free)
        feat[:, idx, :] = P["mask_token"].to(feat.dtype)                        # :373 (in place)
    feat_out = transformer(feat, P, heads)                                      # :375
    feat_out = feat_out.reshape(B, -1)                                          # :377

    pred = mean_params.to(feat_out.dtype).repeat(B, 1).clone()                  # :379-382
    pred[:, 3:] = pred[:, 3:] + feat_out                                        # :383
    for _ in range(iteration):                                                  # :385-387
        out = F.linear(torch.cat((main_feat, pred), dim=1), P["regressor.weight"], P["regressor.bias"])
        pred = pred + out
    pred_3d = pred[:, 3:66].view(-1, 21, 3)                                     # :389
    root = pred_3d[:, 1].clone().unsqueeze(1)
    pred_3d -= root                                                             # :391 (in place on the view)
    pred[:, 3:] = pred_3d.view(-1, 63)                                          # :393
    if pl_reg:
        pl = torch.autograd.grad(torch.sum(feat_out), feat_visual, retain_graph=True)[0]   # :396
        return pred, feat_visual, pl
    return pred, feat_visual


def coarse_forward(P: Dict[str, torch.Tensor], x2, main_feat, mean_params, *, pos_embed: bool = True,
                   mask_idx: Optional[Sequence[int]] = None, pe: Optional[torch.Tensor] = None, depth: int = 3):
    """EncoderTransformerCoarse.forward after the backbone, without the path-length term: hand_net.py:264-311 with the
    transformer of models/vision_transformer_attn.py:104-113 (8 heads, hand_net.py:236).

    Returns (pred_params[B,66], feat_visual[B,21,28,28], attn[B,8,21,21] of the last layer)."""
    heads = 8
    B = x2.shape[0]
    n_tok = P["conv1x1_channel_reduction.weight"].shape[0]
    feat_visual = F.conv2d(x2, P["conv1x1_channel_reduction.weight"])          # :267
    feat = feat_visual.view(B, n_tok, -1)                                       # :268 (a view)
    if pos_embed:
        if pe is None:
            pe = positional_encoding(n_tok, feat.shape[-1], dtype=feat.dtype).to(feat.device)
        feat = feat + pe[: feat.size(0), :]                                     # :272
    if mask_idx is not None and len(mask_idx) > 0:
        feat[:, list(mask_idx), :] = P["mask_token"].to(feat.dtype)             # :280 (in place)
    x, attn = feat, None
    for i in range(depth):                                                      # vision_transformer_attn.py:105-112
        p = f"transformer.layers.{i}."
        last = i == depth - 1
        b, n, _ = x.shape
        qkv = F.linear(x, P[p + "0.to_qkv.weight"])                             # attention on the RAW tokens (:106, :63)
        q, k, v = (t.reshape(b, n, heads, DIM_HEAD).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
        attn = (torch.matmul(q, k.transpose(-1, -2)) * (DIM_HEAD ** -0.5)).softmax(dim=-1)     # :66,76
        o = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, heads * DIM_HEAD)          # :78-79
        x1 = F.linear(o, P[p + "0.to_out.0.weight"], P[p + "0.to_out.0.bias"])                  # :80
        x = F.layer_norm(x1, x1.shape[-1:], P[p + "1.norm.weight"], P[p + "1.norm.bias"], 1e-5) + x   # :107
        if not last:                                                            # :108, PreNorm(FeedForward)
            y = F.layer_norm(x, x.shape[-1:], P[p + "2.norm.weight"], P[p + "2.norm.bias"], 1e-5)
            ff = p + "2.fn.net."
        else:
            y = x
            ff = p + "2.net."
        x = F.linear(F.gelu(F.linear(y, P[ff + "0.weight"], P[ff + "0.bias"])), P[ff + "2.weight"], P[ff + "2.bias"])
    feat_out = x.reshape(B, -1)                                                 # hand_net.py:287
    pred = mean_params.to(feat_out.dtype).repeat(B, 1).clone()                  # :289-293
    pred[:, 3:] = pred[:, 3:] + feat_out                                        # :294
    cameras = F.linear(torch.cat((main_feat, pred[:, :3]), dim=1), P["regressor.weight"], P["regressor.bias"])   # :296
    pred_3d = pred[:, 3:66].view(-1, 21, 3)                                     # :298
    root = pred_3d[:, 1].clone().unsqueeze(1)
    pred_3d -= root                                                             # :300
    pred[:, 3:] = pred_3d.view(-1, 63)
    pred[:, :3] = cameras                                                       # :303
    return pred, feat_visual, attn


def train_loss(pred_params, labels, pl_term=None, *, l_weight_3d: float = 1e5, l_weight_2d: float = 10.0):
    """train.py:165-203 restated (train.py itself is unimportable offline: matplotlib/oss2/...).

    Returns (loss, l_3d, l_2d, l_pl)."""
    cam = pred_params[:, :3]                                                    # :165-167
    j3d = pred_params[:, 3:66].view(-1, 21, 3)
    camera = cam.view(-1, 1, 3)                                                 # :112-118
    x_trans = j3d[:, :, :2] + camera[:, :, 1:]
    res = camera[:, :, 0] * x_trans.reshape(x_trans.size(0), -1)
    j2d = res.view(x_trans.size(0), x_trans.size(1), -1) * 112 + 112            # :119-120
    j3d = j3d.reshape(-1, 63)
    j2d = j2d.reshape(-1, 42)
    if pl_term is not None:                                                     # :178-183
        pl_lengths = torch.sum(torch.square(pl_term), dim=[2, 3]).mean(dim=[1]).sqrt()
        pl_mean = 0.01 * torch.mean(pl_lengths)
        l_pl = torch.square(pl_lengths - pl_mean).mean()
    else:
        l_pl = torch.zeros((), dtype=pred_params.dtype)
    if labels.shape[1] == 105:                                                  # :188-190  MTC / RHD / STB rows
        gt3d, gt2d = labels[:, :63], labels[:, 63:]
    else:                                                                       # :193-196  FreiHAND / HO-3D rows: pose first
        gt3d, gt2d = labels[:, 61:61 + 63], labels[:, 61 + 63:]
    l_3d = F.mse_loss(j3d, gt3d)
    l_2d = F.l1_loss(j2d, gt2d)
    loss = l_weight_3d * l_3d + l_weight_2d * l_2d + (10 * l_pl if pl_term is not None else 0.0)
    return loss, l_3d, l_2d, l_pl


def train_step(P: Dict[str, torch.Tensor], x2, main_feat, labels, mean_params, *, heads=8, iteration=3,
               pos_embed=True, mask_idx=None, pl_reg=True, l_weight_3d=1e5, l_weight_2d=10.0, pe=None):
    """One head training step body (train.py:159-206): forward, path-length VJP, losses, backward.

    Returns dict(loss, l_3d, l_2d, l_pl, pred, feat_visual, pl, grads{name: tensor}, x2_grad, main_feat_grad)."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    x2 = x2.detach().clone().requires_grad_(True)
    main_feat = main_feat.detach().clone().requires_grad_(True)
    outs = head_forward(Pg, x2, main_feat, mean_params, heads=heads, iteration=iteration,
                        pos_embed=pos_embed, mask_idx=mask_idx, pl_reg=pl_reg, pe=pe)
    pred, feat_visual = outs[0], outs[1]
    pl = outs[2] if pl_reg else None
    loss, l3, l2, lpl = train_loss(pred, labels, pl, l_weight_3d=l_weight_3d, l_weight_2d=l_weight_2d)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return dict(loss=loss.detach(), l_3d=l3.detach(), l_2d=l2.detach(), l_pl=lpl.detach(), pred=pred.detach(),
                feat_visual=feat_visual.detach(), pl=None if pl is None else pl.detach(), grads=grads,
                x2_grad=x2.grad, main_feat_grad=main_feat.grad)


# --------------------------------------------------------------------------------------------
# adjacent components (config 4 token transformer, config 5 regressor shapes)
# --------------------------------------------------------------------------------------------
def token_transformer_forward(P, tokens, *, heads=8, pos_embed=True, mask_idx=None, pe=None, depth=3):
    """HRNet-variant token path up to feat.mean(dim=1): hand_net.py:193-203.

    tokens[B,n,dim] -> (feat_out[B,n,3], mean[B,3]).  Beyond the mean the reference raises."""
    feat = tokens
    if pos_embed:
        if pe is None:
            pe = positional_encoding(tokens.shape[1], tokens.shape[2], dtype=tokens.dtype)
        feat = feat + pe[: feat.size(0), :]
    if mask_idx is not None and len(mask_idx) > 0:
        feat = feat.clone()
        feat[:, list(mask_idx), :] = P["mask_token"].to(feat.dtype)
    out = transformer(feat, P, heads, depth)
    return out, out.mean(dim=1)


def h3dw_regressor(main_feat, mean_params, fc2_w, fc2_b, reg_w, reg_b, iters: int = 3):
    """H3DWEncoder.forward after the backbone: hand_net.py:49-58."""
    feat = F.relu(F.linear(F.relu(main_feat), fc2_w, fc2_b))
    pred = mean_params.to(feat.dtype).expand(feat.shape[0], -1)
    for _ in range(iters):
        pred = pred + F.linear(torch.cat([feat, pred], dim=1), reg_w, reg_b)
    return feat, pred


def iterative_regressor(main_feat, pred0, reg_w, reg_b, iters: int = 3):
    """The bare autoregressive loop hand_net.py:385-387 (EncoderTransformer shape 1090->66)."""
    pred = pred0
    for _ in range(iters):
        pred = pred + F.linear(torch.cat([main_feat, pred], dim=1), reg_w, reg_b)
    return pred
