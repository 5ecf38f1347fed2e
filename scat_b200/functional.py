"""Host-side operators over the C ABI (include/scat_b200.h): torch tensors in, torch tensors out.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP on the head's path is
executed by the hand-written sm_100a kernels in scat_b200/csrc.  There is no eager fallback: a missing
library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import PREC, EPI, X2_DTYPE, ScatHeadDesc, check, ptr, ptr_array, stream_ptr


@dataclass(frozen=True)
class HeadConfig:
    heads: int = 8
    iteration: int = 3
    pos_embed: bool = True
    n_masked: int = 0
    pl_reg: bool = False
    precision: str = "tf32"
    n_tokens: int = 21
    channels: int = 512
    token_dim: int = 784
    main_feat_dim: int = 1024
    n_out: int = 66
    x2_dtype: str = "fp32"          # seam storage of x2 and x2.grad: "fp32" (resnet.py:151) or "bf16" (autocast backbone)

    def desc(self, batch: int) -> ScatHeadDesc:
        return ScatHeadDesc(batch, self.n_tokens, self.channels, self.token_dim, self.heads, self.iteration,
                            int(self.pos_embed), self.n_masked, int(self.pl_reg), PREC[self.precision],
                            self.main_feat_dim, self.n_out, X2_DTYPE[self.x2_dtype])


def _seam(t: torch.Tensor, name: str):
    """The backbone seam tensor x2: fp32 or bf16, NCHW contiguous.  Returns (tensor, "fp32" | "bf16")."""
    if not t.is_cuda:
        raise RuntimeError(f"scat_b200: {name} must be a CUDA tensor (the head has no CPU path)")
    if t.dtype == torch.float32:
        return t.contiguous(), "fp32"
    if t.dtype == torch.bfloat16:
        return t.contiguous(), "bf16"
    raise RuntimeError(f"scat_b200: {name} must be float32 or bfloat16, got {t.dtype}")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"scat_b200: {name} must be a CUDA tensor (the head has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"scat_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def workspace_bytes(cfg: HeadConfig, batch: int) -> int:
    d = cfg.desc(batch)
    n = _lib.load().scat_head_workspace_bytes(C.byref(d))
    if n == 0:
        raise RuntimeError("scat_b200: unsupported head configuration: "
                           + _lib.load().scat_last_error_string().decode())
    return n


def alloc_workspace(cfg: HeadConfig, batch: int, device) -> torch.Tensor:
    return torch.empty(workspace_bytes(cfg, batch), dtype=torch.uint8, device=device)


class HeadFunction(torch.autograd.Function):
    """EncoderTransformer.forward after the backbone (hand_net.py:363-398) with a hand-written backward."""

    @staticmethod
    def forward(ctx, cfg: HeadConfig, mask_idx, mean_params, pe, x2, main_feat, *params):
        lib = _lib.load()
        x2, seam = _seam(x2, "x2")
        if seam != cfg.x2_dtype:
            cfg = dataclasses.replace(cfg, x2_dtype=seam)      # the seam dtype follows the tensor the backbone delivered
        if seam == "bf16" and cfg.precision == "fp32":
            raise RuntimeError("scat_b200: a bfloat16 x2 needs precision 'tf32' or 'bf16' (fp32 is the CUDA-core parity mode)")
        main_feat = _f32c(main_feat, "main_feat")
        B = x2.shape[0]
        params = [_f32c(p.detach(), "parameter") for p in params]
        dev = x2.device
        pred = torch.empty(B, cfg.n_out, device=dev, dtype=torch.float32)
        fv = torch.empty(B, cfg.n_tokens, 28, 28, device=dev, dtype=torch.float32)
        pl = torch.empty_like(fv) if cfg.pl_reg else None
        ws = alloc_workspace(cfg, B, dev)
        d = cfg.desc(B)
        check(lib.scat_head_forward(C.byref(d), ptr_array(params), ptr(pe), ptr(mean_params), ptr(mask_idx),
                                    ptr(x2), ptr(main_feat), ptr(pred), ptr(fv), ptr(pl), ptr(ws), ws.numel(),
                                    stream_ptr()), "scat_head_forward")
        ctx.cfg = cfg
        ctx.mask_idx = mask_idx
        ctx.ws = ws
        ctx.set_materialize_grads(False)     # unused outputs arrive as None instead of dense zeros
        ctx.save_for_backward(x2, main_feat, fv, *params)
        ctx.mark_non_differentiable(*([pl] if pl is not None else []))
        return (pred, fv, pl) if cfg.pl_reg else (pred, fv)

    @staticmethod
    def backward(ctx, g_pred, g_fv, *rest):
        lib = _lib.load()
        cfg = ctx.cfg
        x2, main_feat, fv, *params = ctx.saved_tensors
        B = x2.shape[0]
        g_pred = torch.zeros(B, cfg.n_out, device=x2.device) if g_pred is None else _f32c(g_pred, "grad_pred")
        g_fv = None if g_fv is None else _f32c(g_fv, "grad_feat_visual")
        grads = [torch.empty_like(p) for p in params]
        need_x2, need_mf = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        x2_grad = torch.empty_like(x2) if need_x2 else None
        mf_grad = torch.empty_like(main_feat) if need_mf else None
        d = cfg.desc(B)
        check(lib.scat_head_backward(C.byref(d), ptr_array(params), ptr(ctx.mask_idx), ptr(x2), ptr(main_feat),
                                     ptr(fv), ptr(g_pred), ptr(g_fv), ptr_array(grads), ptr(x2_grad), ptr(mf_grad),
                                     ptr(ctx.ws), ctx.ws.numel(), stream_ptr()), "scat_head_backward")
        ctx.ws = None
        return (None, None, None, None, x2_grad, mf_grad, *grads)


def coarse_forward(cfg: HeadConfig, mask_idx, mean_params, pe, x2, main_feat, params):
    """EncoderTransformerCoarse.forward after the backbone (hand_net.py:264-311), inference: returns
    (pred_params[B,66], feat_visual[B,21,28,28], attn[B,heads,21,21])."""
    lib = _lib.load()
    x2, seam = _seam(x2, "x2")
    if seam != cfg.x2_dtype:
        cfg = dataclasses.replace(cfg, x2_dtype=seam)
    main_feat = _f32c(main_feat, "main_feat")
    params = [_f32c(p, "parameter") for p in params]
    B, dev = x2.shape[0], x2.device
    pred = torch.empty(B, cfg.n_out, device=dev, dtype=torch.float32)
    fv = torch.empty(B, cfg.n_tokens, 28, 28, device=dev, dtype=torch.float32)
    attn = torch.empty(B, cfg.heads, cfg.n_tokens, cfg.n_tokens, device=dev, dtype=torch.float32)
    ws = alloc_workspace(cfg, B, dev)
    d = cfg.desc(B)
    check(lib.scat_coarse_forward(C.byref(d), ptr_array(params), ptr(pe), ptr(mean_params), ptr(mask_idx), ptr(x2),
                                  ptr(main_feat), ptr(pred), ptr(fv), ptr(attn), ptr(ws), ws.numel(), stream_ptr()),
          "scat_coarse_forward")
    return pred, fv, attn


class ProjLossFunction(torch.autograd.Function):
    """train.py:112-120,165-203: weak-perspective projection + MSE-3D + L1-2D + path-length statistic."""

    @staticmethod
    def forward(ctx, pred, labels, pl_term, w3d, w2d):
        lib = _lib.load()
        pred = _f32c(pred, "pred_params")
        labels = _f32c(labels, "labels")
        pl = None if pl_term is None else _f32c(pl_term, "pl_term")
        B = pred.shape[0]
        losses = torch.empty(4, device=pred.device, dtype=torch.float32)
        g = torch.empty_like(pred)
        scratch = torch.empty(max(B, 1), device=pred.device, dtype=torch.float32)
        n_tok = pl.shape[1] if pl is not None else 21
        tdim = pl[0, 0].numel() if pl is not None else 784
        check(lib.scat_proj_loss(B, n_tok, tdim, ptr(pred), ptr(labels), labels.shape[1], ptr(pl), float(w3d),
                                 float(w2d), 1.0, ptr(losses), ptr(g), ptr(scratch), stream_ptr()), "scat_proj_loss")
        ctx.save_for_backward(g)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_loss, _g_losses):
        (g,) = ctx.saved_tensors
        return g * g_loss, None, None, None, None


def proj_loss(pred_params, labels, pl_term=None, l_weight_3d=1e5, l_weight_2d=10.0):
    """Returns (loss, [loss, l_3d, l_2d, l_pl]); loss is differentiable w.r.t. pred_params."""
    return ProjLossFunction.apply(pred_params, labels, pl_term, l_weight_3d, l_weight_2d)


def token_transformer(transformer, x, mask_token=None, pe=None, mask_idx=None, precision="tf32", return_mean=False):
    """Inference-only token path (config 4; hand_net.py:193-203): x[B,n,dim] -> [B,n,3]."""
    lib = _lib.load()
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in transformer.parameters())):
        raise RuntimeError("scat_b200.token_transformer is the inference path; wrap the call in torch.no_grad()")
    x = _f32c(x, "tokens")
    B, n, dim = x.shape
    n_masked = 0 if mask_idx is None else int(mask_idx.numel())
    cfg = HeadConfig(heads=transformer.heads, iteration=0, pos_embed=pe is not None, n_masked=n_masked, pl_reg=False,
                     precision=precision, n_tokens=n, channels=0, token_dim=dim, main_feat_dim=0, n_out=0)
    tparams = [p.detach().contiguous() for p in transformer.ordered_parameters()]
    mt = torch.zeros(dim, device=x.device) if mask_token is None else _f32c(mask_token.detach(), "mask_token")
    params = [mt, None] + tparams + [None, None]
    out = torch.empty(B, n, 3, device=x.device, dtype=torch.float32)
    mean = torch.empty(B, 3, device=x.device, dtype=torch.float32)
    ws = alloc_workspace(cfg, B, x.device)
    d = cfg.desc(B)
    check(lib.scat_tokens_forward(C.byref(d), ptr_array(params), ptr(pe), ptr(mask_idx), ptr(x), ptr(out), ptr(mean),
                                  ptr(ws), ws.numel(), stream_ptr()), "scat_tokens_forward")
    return (out, mean) if return_mean else out


# ---------------------------------------------------------------------------------------------------
# single operators (unit tests, micro-benchmarks)
# ---------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_strides=None, b_strides=None, m=None, n=None, k=None, epilogue="none", bias=None, aux_in=None,
         precision="fp32", out=None, accumulate=False, prerounded=False, split_k=False):
    """C[M,N] = epilogue(sum_k A(m,k) B(n,k)).  Defaults: a[M,K] row-major, b[N,K] row-major (y = x W^T).
    Returns C, or (C, Z) for the bias_gelu epilogue."""
    lib = _lib.load()
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if a_strides is None:
        m, k = a.shape
        a_strides = (a.stride(0), a.stride(1))
    if b_strides is None:
        n = b.shape[0]
        b_strides = (b.stride(0), b.stride(1))
    if out is None:
        c = (torch.zeros if split_k else torch.empty)(m, n, device=a.device, dtype=torch.float32)
    else:
        c = out
    z = torch.empty_like(c) if epilogue == "bias_gelu" else None
    check(lib.scat_gemm(ptr(a), a_strides[0], a_strides[1], ptr(b), b_strides[0], b_strides[1], ptr(c), c.stride(0),
                        m, n, k, EPI[epilogue], ptr(bias), ptr(aux_in), 0 if aux_in is None else aux_in.stride(0),
                        ptr(z), 0 if z is None else z.stride(0), int(accumulate),
                        PREC[precision] | (0x100 if prerounded else 0) | (0x200 if split_k else 0), stream_ptr()),
          "scat_gemm")
    return (c, z) if z is not None else c


def gemm_bf16(a, b, *, a_strides=None, b_strides=None, m=None, n=None, k=None, epilogue="none", bias=None,
              aux_in=None, out=None, out16=None, want16=False, split_k=False):
    """bf16-operand tensor-core GEMM (tcgen05 kind::f16): a, b are torch.bfloat16 CUDA tensors; returns the fp32
    result (and the bf16 copy when want16 / out16).  Same stride conventions as gemm()."""
    lib = _lib.load()
    for t, name in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.bfloat16:
            raise RuntimeError(f"scat_b200: {name} must be a CUDA bfloat16 tensor")
    a, b = a.contiguous(), b.contiguous()
    if a_strides is None:
        m, k = a.shape
        a_strides = (a.stride(0), a.stride(1))
    if b_strides is None:
        n = b.shape[0]
        b_strides = (b.stride(0), b.stride(1))
    if out is None:
        c = torch.zeros(m, n, device=a.device, dtype=torch.float32) if split_k else torch.empty(m, n, device=a.device, dtype=torch.float32)
    else:
        c = out
    if out16 is None and want16:
        out16 = torch.empty(m, (n + 7) // 8 * 8, device=a.device, dtype=torch.bfloat16)
    z = torch.empty_like(c) if epilogue == "bias_gelu" else None
    check(lib.scat_gemm_bf16(ptr(a), a_strides[0], a_strides[1], ptr(b), b_strides[0], b_strides[1], ptr(c), c.stride(0),
                             ptr(out16), 0 if out16 is None else out16.stride(0), m, n, k, EPI[epilogue], ptr(bias),
                             ptr(aux_in), 0 if aux_in is None else aux_in.stride(0), ptr(z),
                             0 if z is None else z.stride(0), int(split_k), stream_ptr()), "scat_gemm_bf16")
    res = [c]
    if z is not None:
        res.append(z)
    if out16 is not None:
        res.append(out16)
    return res[0] if len(res) == 1 else tuple(res)


def layernorm_fwd(x, gamma, beta):
    lib = _lib.load()
    x = _f32c(x, "x")
    rows, dim = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device)
    rstd = torch.empty(rows, device=x.device)
    check(lib.scat_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, dim, stream_ptr()),
          "scat_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, resid=None, param_grads=True):
    lib = _lib.load()
    rows, dim = x.shape
    dx = torch.empty_like(x)
    dg = torch.empty(dim, device=x.device) if param_grads else None
    db = torch.empty(dim, device=x.device) if param_grads else None
    check(lib.scat_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(resid), ptr(dx), ptr(dg),
                                 ptr(db), rows, dim, stream_ptr()), "scat_layernorm_bwd")
    return dx, dg, db


def attention_fwd(qkv, batch, n, heads, tc=False):
    """vision_transformer.py:61-77.  tc=True (n = 21 only): the mma.sync TF32 kernel of the TF32 / BF16 precisions."""
    lib = _lib.load()
    qkv = _f32c(qkv, "qkv")
    o = torch.empty(batch * n, heads * 64, device=qkv.device)
    p = torch.empty(batch, heads, n, n, device=qkv.device)
    fn = lib.scat_attention_fwd_tc if tc else lib.scat_attention_fwd
    check(fn(ptr(qkv), ptr(o), ptr(p), batch, n, heads, stream_ptr()), "scat_attention_fwd")
    return o, p


def attention_bwd(qkv, p, d_o, batch, n, heads, tc=False):
    lib = _lib.load()
    dqkv = torch.empty_like(qkv)
    fn = lib.scat_attention_bwd_tc if tc else lib.scat_attention_bwd
    check(fn(ptr(qkv), ptr(p), ptr(_f32c(d_o, "d_o")), ptr(dqkv), batch, n, heads, stream_ptr()), "scat_attention_bwd")
    return dqkv


def conv_pe_mask_fwd(x2, conv_w, pe, mask_token, mask_idx, pos_embed=True, tc=False):
    """hand_net.py:363-373.  tc=False: fp32 FFMA kernel (parity mode); tc=True: the persistent tcgen05 kernel
    (csrc/conv_tc.cu), x2 fp32 (kind::tf32) or bfloat16 (kind::f16)."""
    lib = _lib.load()
    x2, seam = _seam(x2, "x2")
    if seam == "bf16" and not tc:
        raise RuntimeError("scat_b200: a bfloat16 x2 needs the tensor-core front end (tc=True)")
    B, Cc, H, W = x2.shape
    T = conv_w.shape[0]
    fv = torch.empty(B, T, H, W, device=x2.device, dtype=torch.float32)
    tok = torch.empty(B, T, H * W, device=x2.device, dtype=torch.float32) if pos_embed else fv
    n_masked = 0 if mask_idx is None else int(mask_idx.numel())
    if tc:
        scratch = torch.empty(lib.scat_conv_tc_scratch_floats(B, Cc, H * W, T), device=x2.device)
        check(lib.scat_conv_pe_mask_fwd_tc(ptr(x2), X2_DTYPE[seam], ptr(_f32c(conv_w, "conv_w")), ptr(pe), ptr(mask_token), ptr(mask_idx),
                                           n_masked, int(pos_embed), ptr(fv), ptr(tok), ptr(scratch), B, Cc, H * W, T,
                                           stream_ptr()), "scat_conv_pe_mask_fwd_tc")
        return fv, tok.view(B, T, H * W)
    check(lib.scat_conv_pe_mask_fwd(ptr(x2), ptr(_f32c(conv_w, "conv_w")), ptr(pe), ptr(mask_token), ptr(mask_idx),
                                    n_masked, int(pos_embed), ptr(fv), ptr(tok), B, Cc, H * W, T, stream_ptr()),
          "scat_conv_pe_mask_fwd")
    return fv, tok.view(B, T, H * W)


def conv_bwd(d_tokens, x2, conv_w, mask_idx, need_x2_grad=True, tc=False):
    """Backward of the conv front end; x2.grad comes back in x2's dtype (fp32, or bfloat16 with tc=True)."""
    lib = _lib.load()
    x2, seam = _seam(x2, "x2")
    if seam == "bf16" and not tc:
        raise RuntimeError("scat_b200: a bfloat16 x2 needs the tensor-core front end (tc=True)")
    B, Cc, H, W = x2.shape
    T = conv_w.shape[0]
    n_masked = 0 if mask_idx is None else int(mask_idx.numel())
    x2g = torch.empty_like(x2) if need_x2_grad else None
    wg = torch.empty(T, Cc, device=x2.device, dtype=torch.float32)
    mg = torch.zeros(H * W, device=x2.device)
    if tc:
        scratch = torch.empty(lib.scat_conv_tc_scratch_floats(B, Cc, H * W, T), device=x2.device)
        check(lib.scat_conv_bwd_tc(ptr(_f32c(d_tokens, "d_tokens")), ptr(x2), X2_DTYPE[seam], ptr(_f32c(conv_w, "conv_w")), ptr(mask_idx),
                                   n_masked, ptr(x2g), ptr(wg), ptr(mg) if n_masked else None, ptr(scratch), B, Cc, H * W,
                                   T, stream_ptr()), "scat_conv_bwd_tc")
        return x2g, wg, mg
    scratch = torch.empty(lib.scat_conv_bwd_scratch_floats(B, Cc, H * W, T), device=x2.device)
    check(lib.scat_conv_bwd(ptr(_f32c(d_tokens, "d_tokens")), ptr(x2), ptr(_f32c(conv_w, "conv_w")), ptr(mask_idx),
                            n_masked, ptr(x2g), ptr(wg), ptr(mg) if n_masked else None, ptr(scratch), B, Cc, H * W, T,
                            stream_ptr()), "scat_conv_bwd")
    return x2g, wg, mg


def regressor_fwd(main_feat, feat_out, mean_params, w, b, iteration=3, root_relative=True, keep_states=False):
    """hand_net.py:379-393 (root_relative) / hand_net.py:53-57 (H3DWEncoder: feat_out=None, root_relative=False)."""
    lib = _lib.load()
    main_feat = _f32c(main_feat, "main_feat")
    B, F = main_feat.shape
    P = w.shape[0]
    pred = torch.empty(B, P, device=main_feat.device)
    states = torch.empty(B, max(iteration, 1), P, device=main_feat.device) if keep_states else None
    scratch = torch.empty(B, P, device=main_feat.device)
    check(lib.scat_regressor_fwd(ptr(main_feat), ptr(feat_out), ptr(_f32c(mean_params, "mean_params")),
                                 ptr(_f32c(w, "w")), ptr(_f32c(b, "b")), ptr(pred), ptr(states), ptr(scratch), B, F, P,
                                 iteration, int(root_relative), stream_ptr()), "scat_regressor_fwd")
    return (pred, states) if keep_states else pred
