"""Deterministic synthetic inputs, weights and assets for the reg_transformer head.

The reference's datasets (STB / HO-3D / FreiHAND) and MANO_RIGHT.pkl are not
available offline, so every test, the oracle pinning script
and bench.py draw their tensors from this module.  Everything is generated with
numpy's PCG64 bit generator so the same seed gives the same bytes on the
authoring container and on the GPU box, independent of the torch build.

Shapes follow the reference:
  * head parameters        models/hand_net.py:319-353, models/vision_transformer.py:81-96
  * backbone seam tensors  models/resnet.py:142-162 (x2 [B,512,28,28], main_feat [B,1024], both post-ReLU)
  * labels                 dataset/load_STB.py:286-295 (63 root-relative 3D metres + 42 2D pixels)
  * MANO-shaped asset      models/mano.py:215-234
"""
from __future__ import annotations

import numpy as np

N_TOKENS = 21          # hand_net.py:328  full_content
TOKEN_DIM = 784        # 28*28 flattened spatial map, hand_net.py:331
X2_CHANNELS = 512      # resnet layer2 output, hand_net.py:329
MAIN_FEAT = 1024       # resnet fc1 output, resnet.py:159
N_PARAMS_OUT = 66      # 3 camera + 21*3 joints, hand_net.py:353
DIM_HEAD = 64          # hand_net.py:331


def layer_dims(dim: int = TOKEN_DIM, depth: int = 3):
    """(dim, hidden, out) per transformer layer; vision_transformer.py:84-96."""
    dims = []
    for i in range(depth):
        hid = (dim * 3) // 4
        out = 3 if i == depth - 1 else dim // 2
        dims.append((dim, hid, out))
        dim = dim // 2
    return dims


def head_param_shapes(heads: int = 8, dim: int = TOKEN_DIM, depth: int = 3, channels: int = X2_CHANNELS,
                      n_tokens: int = N_TOKENS):
    """Ordered {state_dict key: shape} for the non-backbone parameters of EncoderTransformer.

    Order is the reference's named_parameters() order (hand_net.py:329-353)."""
    inner = DIM_HEAD * heads
    shapes = {}
    shapes["mask_token"] = (1, 1, dim)
    shapes["conv1x1_channel_reduction.weight"] = (n_tokens, channels, 1, 1)
    for i, (d, hid, out) in enumerate(layer_dims(dim, depth)):
        p = f"transformer.layers.{i}."
        shapes[p + "0.fn.norm.weight"] = (d,)
        shapes[p + "0.fn.norm.bias"] = (d,)
        shapes[p + "0.fn.fn.to_qkv.weight"] = (3 * inner, d)
        shapes[p + "0.fn.fn.to_out.0.weight"] = (d, inner)
        shapes[p + "0.fn.fn.to_out.0.bias"] = (d,)
        if i < depth - 1:
            shapes[p + "1.norm.weight"] = (d,)
            shapes[p + "1.norm.bias"] = (d,)
            ff = p + "1.fn.net."
        else:
            ff = p + "1.net."
        shapes[ff + "0.weight"] = (hid, d)
        shapes[ff + "0.bias"] = (hid,)
        shapes[ff + "2.weight"] = (out, hid)
        shapes[ff + "2.bias"] = (out,)
    shapes["regressor.weight"] = (N_PARAMS_OUT, MAIN_FEAT + N_PARAMS_OUT)
    shapes["regressor.bias"] = (N_PARAMS_OUT,)
    return shapes


def coarse_param_shapes(dim: int = TOKEN_DIM, depth: int = 3, channels: int = X2_CHANNELS, n_tokens: int = N_TOKENS):
    """Ordered {state_dict key: shape} of EncoderTransformerCoarse (hand_net.py:220-259, vision_transformer_attn.py:88-104):
    8 heads fixed, LayerNorm on the attention output (layers.i.1), camera regressor Linear(1027 -> 3)."""
    inner = DIM_HEAD * 8
    shapes = {"mask_token": (1, 1, dim), "conv1x1_channel_reduction.weight": (n_tokens, channels, 1, 1)}
    for i, (d, hid, out) in enumerate(layer_dims(dim, depth)):
        p = f"transformer.layers.{i}."
        shapes[p + "0.to_qkv.weight"] = (3 * inner, d)
        shapes[p + "0.to_out.0.weight"] = (d, inner)
        shapes[p + "0.to_out.0.bias"] = (d,)
        shapes[p + "1.norm.weight"] = (d,)
        shapes[p + "1.norm.bias"] = (d,)
        if i < depth - 1:
            shapes[p + "2.norm.weight"] = (d,)
            shapes[p + "2.norm.bias"] = (d,)
            ff = p + "2.fn.net."
        else:
            ff = p + "2.net."
        shapes[ff + "0.weight"] = (hid, d)
        shapes[ff + "0.bias"] = (hid,)
        shapes[ff + "2.weight"] = (out, hid)
        shapes[ff + "2.bias"] = (out,)
    shapes["regressor.weight"] = (3, MAIN_FEAT + 3)
    shapes["regressor.bias"] = (3,)
    return shapes


def make_coarse_weights(seed: int = 20210206, dtype=np.float32):
    """Random-init weights of the coarse head, same recipe as make_head_weights("unit")."""
    g = _rng(seed)
    out = {}
    for name, shape in coarse_param_shapes().items():
        r = g.standard_normal(size=shape)
        out[name] = ((1.0 + 0.02 * r) if name.endswith("norm.weight") else 0.02 * r).astype(dtype)
    return out


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def make_head_weights(heads: int = 8, seed: int = 20211011, regime: str = "unit", dtype=np.float32):
    """Random-init head weights.

    regime "unit":  norm.weight = 1 + 0.02 r, everything else 0.02 r (r ~ N(0,1)); outputs are O(1).
    regime "hand":  same, then the last FF output layer and the regressor are scaled so that the
                    predicted joint offsets are ~10 mm around the mean template (metres), which is
                    the regime in which the 0.05 mm MPJPE budget of the bf16 path is meaningful.
    """
    g = _rng(seed)
    out = {}
    for name, shape in head_param_shapes(heads).items():
        r = g.standard_normal(size=shape)
        if name.endswith("norm.weight"):
            w = 1.0 + 0.02 * r
        else:
            w = 0.02 * r
        out[name] = w.astype(dtype)
    if regime == "hand":
        for name in ("transformer.layers.2.1.net.2.weight", "transformer.layers.2.1.net.2.bias"):
            out[name] = (out[name] * 0.05).astype(dtype)
        for name in ("regressor.weight", "regressor.bias"):
            out[name] = (out[name] * 0.01).astype(dtype)
    elif regime != "unit":
        raise ValueError(f"unknown weight regime {regime!r}")
    return out


def make_mean_params(kind: str = "hand", dtype=np.float32):
    """mean_params[1,66]: camera (s,tx,ty) = (5,0,0) then 21 template joints (train.py:94-110).

    "hand": a deterministic hand-like skeleton in metres (wrist at the origin, five fingers of four
    joints fanning out in the x/y plane), standing in for the 21 MANO template vertices that the
    reference reads from MANO_RIGHT.pkl.  "zero": all joints at the origin."""
    mean = np.zeros((1, N_PARAMS_OUT), dtype=np.float64)
    mean[0, 0] = 5.0
    if kind == "hand":
        joints = np.zeros((21, 3))
        # finger k, bone j: train.py's docstring order 13,14,15 / 1,2,3 / 4,5,6 / 10,11,12 / 7,8,9 + tips 16..20
        bases = {1: 0.0, 4: 0.35, 10: 0.7, 7: 1.0, 13: -0.6}
        tips = {1: 16, 4: 17, 10: 19, 7: 18, 13: 20}
        for first, ang in bases.items():
            direction = np.array([np.sin(ang), np.cos(ang), 0.0])
            root = 0.085 * direction + np.array([0.0, 0.0, 0.004 * first])
            for j in range(3):
                joints[first + j] = root + direction * 0.028 * j
            joints[tips[first]] = root + direction * 0.028 * 3
        mean[0, 3:] = joints.reshape(-1)
    elif kind != "zero":
        raise ValueError(f"unknown mean_params kind {kind!r}")
    return mean.astype(dtype)


def make_head_inputs(batch: int, seed: int = 0, channels: int = X2_CHANNELS, hw: int = 28, dtype=np.float32):
    """(x2[B,C,hw,hw], main_feat[B,1024], labels[B,105]) as the backbone/dataset would hand them over."""
    g = _rng(1000003 * (seed + 1))
    x2 = np.maximum(g.standard_normal(size=(batch, channels, hw, hw), dtype=np.float32), 0.0)
    main_feat = np.maximum(g.standard_normal(size=(batch, MAIN_FEAT), dtype=np.float32), 0.0)
    j3d = 0.03 * g.standard_normal(size=(batch, 21, 3))
    j3d -= j3d[:, 1:2, :]                       # labels are root-relative to joint 1 (load_STB.py:286-289)
    j2d = g.uniform(0.0, 224.0, size=(batch, 42))
    labels = np.concatenate([j3d.reshape(batch, 63), j2d], axis=1)
    return x2.astype(dtype), main_feat.astype(dtype), labels.astype(dtype)


def make_token_weights(dim: int, heads: int = 8, seed: int = 99, depth: int = 3, dtype=np.float32):
    """Weights of the bare token transformer (config 4): `transformer.*` keys + mask_token for width `dim`."""
    g = _rng(seed)
    out = {}
    for name, shape in head_param_shapes(heads, dim, depth).items():
        if not (name.startswith("transformer.") or name == "mask_token"):
            continue
        r = g.standard_normal(size=shape)
        out[name] = ((1.0 + 0.02 * r) if name.endswith("norm.weight") else 0.05 * r).astype(dtype)
    return out


def make_token_inputs(batch: int, n_tokens: int, dim: int, seed: int = 0, dtype=np.float32):
    """Config-4 style token input N(0,1)[B,n,dim] (hand_net.py:193-203)."""
    g = _rng(7000001 * (seed + 1))
    return g.standard_normal(size=(batch, n_tokens, dim), dtype=np.float32).astype(dtype)


MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)   # mano.py:221-223 from kintree_table
MANO_TIP_VERTS = (320, 443, 671, 554, 744)                              # mano.py:373-377


def make_mano_asset(seed: int = 7, dtype=np.float32):
    """Random MANO-shaped asset: 778 verts, 16 joints, 10 shape dirs, 135 pose dirs (mano.py:215-234).

    Returns plain numpy arrays keyed as in MANO_RIGHT.pkl; J_regressor is dense [16,778]."""
    g = _rng(seed)
    nv, nj = 778, 16
    v_template = 0.05 * g.standard_normal(size=(nv, 3))
    shapedirs = 0.005 * g.standard_normal(size=(nv, 3, 10))
    posedirs = 0.002 * g.standard_normal(size=(nv, 3, 135))
    jr = g.uniform(0.0, 1.0, size=(nj, nv)) ** 8            # sparse-ish positive rows
    jr[jr < 0.2] = 0.0
    jr[:, :nj] += np.eye(nj)
    jr /= jr.sum(axis=1, keepdims=True)
    w = g.uniform(0.0, 1.0, size=(nv, nj)) ** 6
    w /= w.sum(axis=1, keepdims=True)
    hands_mean = 0.1 * g.standard_normal(size=(45,))
    hands_components = np.eye(45)
    kintree = np.zeros((2, nj), dtype=np.int64)
    kintree[0] = [p if p >= 0 else 2 ** 32 - 1 for p in MANO_PARENTS]
    kintree[1] = np.arange(nj)
    faces = np.stack([np.arange(1538) % nv, (np.arange(1538) + 1) % nv, (np.arange(1538) + 2) % nv], axis=1)
    return {
        "kintree_table": kintree,
        "f": faces.astype(np.int64),
        "v_template": v_template.astype(dtype),
        "shapedirs": shapedirs.astype(dtype),
        "posedirs": posedirs.astype(dtype),
        "J_regressor": jr.astype(dtype),
        "weights": w.astype(dtype),
        "hands_components": hands_components.astype(dtype),
        "hands_mean": hands_mean.astype(dtype),
    }


def make_mano_inputs(batch: int, seed: int = 0, dtype=np.float32):
    """rots ~ N(0,.5^2)[B,3], poses ~ N(0,.3^2)[B,45], betas ~ N(0,1)[B,10]."""
    g = _rng(424243 * (seed + 1))
    rots = 0.5 * g.standard_normal(size=(batch, 3))
    poses = 0.3 * g.standard_normal(size=(batch, 45))
    betas = g.standard_normal(size=(batch, 10))
    return rots.astype(dtype), poses.astype(dtype), betas.astype(dtype)


def mask_indices(mask_rate: float, n_tokens: int = N_TOKENS, rng=None):
    """Host-side token mask draw; hand_net.py:369-372.

    Consumes exactly one ``random.shuffle(list(range(n_tokens)))`` from ``rng`` (default: the global
    ``random`` module, like the reference) when 0.1 <= mask_rate <= 0.9, nothing otherwise."""
    import random as _random
    if not (mask_rate >= 0.1 and mask_rate <= 0.9):
        return []
    masked = list(range(n_tokens))
    (rng or _random).shuffle(masked)
    return masked[: int(mask_rate * n_tokens)]
