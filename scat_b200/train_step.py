"""Head training step (train.py:138-209 body) as one C-ABI call, CUDA-graph replayable, data parallel.

``HeadTrainStep`` owns a flat fp32 parameter-gradient bucket whose slices are the ``.grad`` of the head's
35 tensors, persistent input/output buffers and a caller-visible workspace.  ``step()`` runs

    forward -> path-length VJP -> projection + losses -> backward          (scat_head_train_step)
    -> all-reduce(sum) of the flat gradient bucket over NCCL when world_size > 1

Samples are independent, so data parallelism is pure batch sharding: every rank holds the
same weights and the same host-drawn mask indices, the local loss gradient is scaled by 1/world_size inside
the loss kernel (train.py:191-192 divide by the *global* B*63 / B*42) and the only collective is the gradient
all-reduce.  The optimizer is outside this class (section 8f "next").

Input slots: the step reads its batch from one of ``input_slots`` persistent device buffer sets.  With two slots
a caller can stage batch i+1 from pinned host memory on a copy stream (``load_inputs(..., slot, stream)``) while
batch i computes; ``step(slot)`` orders itself after that slot's copy and the next copy into the slot orders
itself after the step that read it.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib, dp
from . import functional as SF
from ._lib import check, ptr, ptr_array


class HeadTrainStep:
    def __init__(self, net, batch: int, l_weight_3d: float = 1e5, l_weight_2d: float = 10.0, *,
                 need_x2_grad: bool = True, need_main_feat_grad: bool = True, use_graph: bool = True,
                 process_group=None, input_slots: int = 1):
        self.net = net
        self.batch = int(batch)
        self.w3d, self.w2d = float(l_weight_3d), float(l_weight_2d)
        self.lib = _lib.load()
        dev = net.mask_token.device
        if dev.type != "cuda":
            raise RuntimeError("HeadTrainStep: move the module to a CUDA device first (no CPU path)")
        self.device = dev
        self.pg = process_group
        self.world = dp.world_size(process_group)
        self.params = net.head_parameters()
        self.bucket = dp.FlatGradBucket(self.params)      # p.grad are views of one flat buffer
        r = net.mask_rate
        self.n_masked = int(r * net.full_content) if (r >= 0.1 and r <= 0.9) else 0
        self.cfg = net.config(self.n_masked)
        B = self.batch
        self.n_slots = int(input_slots)
        self.x2s = [torch.empty(B, 512, 28, 28, device=dev) for _ in range(self.n_slots)]
        self.main_feats = [torch.empty(B, 1024, device=dev) for _ in range(self.n_slots)]
        self.labelss = [torch.empty(B, 105, device=dev) for _ in range(self.n_slots)]
        self.copied: List[Optional[torch.cuda.Event]] = [None] * self.n_slots     # slot contents are ready
        self.consumed: List[Optional[torch.cuda.Event]] = [None] * self.n_slots   # slot may be overwritten
        self.pred = torch.empty(B, 66, device=dev)
        self.feat_visual = torch.empty(B, 21, 28, 28, device=dev)
        self.pl = torch.empty(B, 21, 28, 28, device=dev) if self.cfg.pl_reg else None
        self.losses = torch.zeros(4, device=dev)
        self.x2_grad = torch.empty(B, 512, 28, 28, device=dev) if need_x2_grad else None
        self.main_feat_grad = torch.empty(B, 1024, device=dev) if need_main_feat_grad else None
        self.mask_dev = torch.zeros(max(self.n_masked, 1), dtype=torch.int32, device=dev)
        self.ws = SF.alloc_workspace(self.cfg, B, dev)
        self.last_mask = []
        self.use_graph = use_graph
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * self.n_slots

    # single-slot views kept for callers that use one buffer set
    @property
    def x2(self):
        return self.x2s[0]

    @property
    def main_feat(self):
        return self.main_feats[0]

    @property
    def labels(self):
        return self.labelss[0]

    # ------------------------------------------------------------------------------------------
    def _enqueue(self, slot: int = 0):
        """Enqueue one fused step on the current stream (graph-capturable: no allocation, no sync)."""
        cfg = self.cfg
        d = cfg.desc(self.batch)
        pe = self.net.positionalEncoding.pe[0] if cfg.pos_embed else None
        labels = self.labelss[slot]
        check(self.lib.scat_head_train_step(
            C.byref(d), ptr_array([p.data for p in self.params]), ptr(pe), ptr(self.net.mean_params.reshape(-1)),
            ptr(self.mask_dev) if self.n_masked else None, ptr(self.x2s[slot]), ptr(self.main_feats[slot]),
            ptr(labels), labels.shape[1], self.w3d, self.w2d, 1.0 / self.world, ptr(self.pred),
            ptr(self.feat_visual), ptr(self.pl), ptr(self.losses), ptr_array(self.bucket.views), ptr(self.x2_grad),
            ptr(self.main_feat_grad), ptr(self.ws), self.ws.numel(), SF.stream_ptr()), "scat_head_train_step")

    def _capture(self, slot: int):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):            # warm-up outside capture: function attributes, module loading
                self._enqueue(slot)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue(slot)
        self.graphs[slot] = g

    def set_mask(self, mask_idx=None):
        """Draw (or take) the token mask on the host and stage it for the next step.  Consumes exactly one
        ``random.shuffle`` like the reference forward (hand_net.py:370-372)."""
        masked = self.net._draw_mask() if mask_idx is None else list(mask_idx)
        if len(masked) != self.n_masked:
            raise ValueError(f"mask has {len(masked)} indices, configuration expects {self.n_masked}")
        self.last_mask = masked
        if self.n_masked:
            self.mask_dev.copy_(torch.tensor(masked, dtype=torch.int32))    # pageable source: staged synchronously
        return masked

    def load_inputs(self, x2, main_feat, labels, slot: int = 0, stream: Optional[torch.cuda.Stream] = None,
                    non_blocking: bool = True):
        """Copy one batch (pinned host or device tensors) into input slot ``slot``; on ``stream`` if given
        (a copy stream), ordered after the last step that read the slot."""
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            cur = torch.cuda.current_stream()
            if self.consumed[slot] is not None:
                cur.wait_event(self.consumed[slot])
            self.x2s[slot].copy_(x2.view_as(self.x2s[slot]), non_blocking=non_blocking)
            self.main_feats[slot].copy_(main_feat, non_blocking=non_blocking)
            self.labelss[slot].copy_(labels[:, :105], non_blocking=non_blocking)
            ev = torch.cuda.Event()
            ev.record(cur)
            self.copied[slot] = ev

    def step(self, allreduce: bool = True, slot: int = 0):
        """Run fwd + pl VJP + loss + bwd on the inputs staged in ``slot``; returns the device tensor losses[4]
        = [loss, l_3d, l_2d, l_pl] (local to this rank's shard)."""
        cur = torch.cuda.current_stream()
        if self.copied[slot] is not None:
            cur.wait_event(self.copied[slot])
        if self.use_graph:
            if self.graphs[slot] is None:
                self._capture(slot)
            self.graphs[slot].replay()
        else:
            self._enqueue(slot)
        if self.n_slots > 1:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.consumed[slot] = ev
        if allreduce:
            self.bucket.all_reduce(self.pg)
        return self.losses


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
