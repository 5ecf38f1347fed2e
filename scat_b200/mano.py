"""Host mirror of the reference's models/mano.py:280-391 over the fused LBS kernel.

``ManoLayer(asset)`` uploads a MANO-shaped asset (keys as in MANO_RIGHT.pkl: v_template, shapedirs, posedirs,
J_regressor (dense or scipy sparse), weights, hands_mean) once, derives the vertex-contiguous tables on the
GPU and exposes ``rot_pose_beta_to_mesh(rots, poses, betas) -> [B, 21+778, 3]`` with the reference's argument
meaning (no PCA: poses are 45 axis-angle values added to hands_mean; local root rotation forced to 0;
joints = 16 chain joints + 5 fingertip vertices; everything relative to joint 1).  Differentiable w.r.t. rots, poses and
betas like the reference (which runs it under autograd): the backward is the ``scat_lbs_bwd`` kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class _LbsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, rots, poses, betas):
        out = layer._forward(rots, poses, betas)
        ctx.layer = layer
        ctx.save_for_backward(rots, poses, betas)
        return out

    @staticmethod
    def backward(ctx, g_out):
        rots, poses, betas = ctx.saved_tensors
        lib = _lib.load()
        B = rots.shape[0]
        g_out = g_out.contiguous().float()
        g_r, g_p, g_b = torch.empty_like(rots), torch.empty_like(poses), torch.empty_like(betas)
        layer = ctx.layer
        check(lib.scat_lbs_bwd(ptr(layer.derived), ptr(layer.hands_mean), ptr(rots), ptr(poses), ptr(betas), ptr(g_out),
                               ptr(g_r), ptr(g_p), ptr(g_b), B, stream_ptr()), "scat_lbs_bwd")
        return None, g_r, g_p, g_b


class ManoLayer:
    """``precision``: "tf32x3" (default) runs the blend-shape contractions as one tcgen05 GEMM at fp32 grade (TF32 hi/lo
    split, csrc/lbs_tc.cu); "fp32" keeps everything on CUDA cores (csrc/lbs.cu).  Both agree to ~1e-7 m."""

    def __init__(self, asset: dict, device=None, precision: str = "tf32x3"):
        if precision not in ("tf32x3", "fp32"):
            raise ValueError("ManoLayer: precision 'tf32x3' or 'fp32'")
        self.precision = precision
        if not torch.cuda.is_available():
            raise RuntimeError("scat_b200.mano needs a CUDA device; there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lib = _lib.load()

        def up(key, shape):
            a = asset[key]
            if hasattr(a, "todense"):
                a = np.asarray(a.todense())
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(self.device)
            if tuple(t.shape) != shape:
                raise ValueError(f"MANO asset {key}: expected {shape}, got {tuple(t.shape)}")
            return t.contiguous()

        self.v_template = up("v_template", (778, 3))
        self.shapedirs = up("shapedirs", (778, 3, 10))
        self.posedirs = up("posedirs", (778, 3, 135))
        self.J_regressor = up("J_regressor", (16, 778))
        self.weights = up("weights", (778, 16))
        self.hands_mean = up("hands_mean", (45,))
        self.derived = torch.empty(lib.scat_lbs_derived_floats(), device=self.device, dtype=torch.float32)
        self.table = torch.empty(lib.scat_lbs_tc_table_floats(), device=self.device, dtype=torch.float32)
        self._scratch = None
        with torch.cuda.device(self.device):
            check(lib.scat_lbs_prepare(ptr(self.v_template), ptr(self.shapedirs), ptr(self.posedirs),
                                       ptr(self.J_regressor), ptr(self.weights), ptr(self.derived), stream_ptr()),
                  "scat_lbs_prepare")
            check(lib.scat_lbs_tc_prepare(ptr(self.shapedirs), ptr(self.posedirs), ptr(self.table), stream_ptr()),
                  "scat_lbs_tc_prepare")

    def _forward(self, rots, poses, betas, out=None):
        lib = _lib.load()
        B = rots.shape[0]
        if out is None:
            out = torch.empty(B, 799, 3, device=rots.device, dtype=torch.float32)
        if self.precision == "fp32":
            check(lib.scat_lbs_fwd(ptr(self.derived), ptr(self.hands_mean), ptr(rots), ptr(poses), ptr(betas), ptr(out), B,
                                   stream_ptr()), "scat_lbs_fwd")
            return out
        need = lib.scat_lbs_tc_scratch_floats(B)
        if self._scratch is None or self._scratch.numel() < need:      # grows to the largest chunk seen, then is reused
            self._scratch = torch.empty(need, device=self.device, dtype=torch.float32)
        check(lib.scat_lbs_fwd_tc(ptr(self.derived), ptr(self.table), ptr(self.hands_mean), ptr(rots), ptr(poses), ptr(betas),
                                  ptr(out), B, ptr(self._scratch), self._scratch.numel(), stream_ptr()), "scat_lbs_fwd_tc")
        return out

    def rot_pose_beta_to_mesh(self, rots, poses, betas, out=None):
        for name, t, w in (("rots", rots, 3), ("poses", poses, 45), ("betas", betas, 10)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == w):
                raise RuntimeError(f"scat_b200.mano: {name} must be a CUDA float32 [B,{w}] tensor")
        rots, poses, betas = rots.contiguous(), poses.contiguous(), betas.contiguous()
        if torch.is_grad_enabled() and (rots.requires_grad or poses.requires_grad or betas.requires_grad):
            if out is not None:
                raise RuntimeError("scat_b200.mano: out= cannot be combined with autograd")
            return _LbsFunction.apply(self, rots, poses, betas)
        return self._forward(rots, poses, betas, out)

    __call__ = rot_pose_beta_to_mesh
