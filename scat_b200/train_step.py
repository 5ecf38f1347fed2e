"""Head training step (train.py:138-209 body) as one C-ABI call, CUDA-graph replayable, data parallel.

``HeadTrainStep`` owns a flat fp32 parameter-gradient bucket whose slices are the ``.grad`` of the head's
35 tensors, persistent input/output buffers and a caller-visible workspace.  ``step()`` runs

    forward -> path-length VJP -> projection + losses -> backward          (scat_head_train_step)
    -> all-reduce(sum) of the flat gradient bucket when world_size > 1

The all-reduce (``comm``): "peer" (default on GPUs of one node) keeps the bucket in CUDA-IPC peer-mapped memory and
sums it with ONE kernel of this library over NVLink (dp.PeerMemory, csrc/dp_allreduce.cu), captured into the same CUDA
graph as the step -- no host work, no extra launch gap, bit-identical sums on every rank.  "nccl" calls
``torch.distributed.all_reduce`` after the step; with ``phased=True`` the step is issued in three phases
(scat_head_train_step_phase) and the all-reduce of each finished part of the bucket runs on a communication stream
under the next phase (measured on 8 B200s this does not beat the single NCCL call, profiles/README.md, so it is
opt-in).

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
    _RING = 8            # pinned staging slots for the mask indices
    POLL_EVERY = 256     # steps between looks at the gradient exchange's time-out flag (each look drains the device)

    def __init__(self, net, batch: int, l_weight_3d: float = 1e5, l_weight_2d: float = 10.0, *,
                 need_x2_grad: bool = True, need_main_feat_grad: bool = True, use_graph: bool = True,
                 process_group=None, input_slots: int = 1, phased: Optional[bool] = None,
                 comm: str = "auto", x2_dtype: str = "fp32", label_width: int = 105, overlap_exchange: bool = True):
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
        self.flat_params = dp.flatten_parameters(self.params)   # one buffer: the fused optimiser walks it linearly
        self.opt = None
        if comm not in ("auto", "peer", "nccl"):
            raise ValueError(f"comm={comm!r}: 'auto', 'peer' or 'nccl'")
        auto = comm == "auto"
        if auto:
            comm = "peer" if self.world in (2, 4, 8) else "nccl"
        self.comm = comm if self.world > 1 else "none"
        self.peer = None
        if self.comm == "peer":                            # bucket lives in memory every rank of the node has mapped
            try:
                self.peer = dp.PeerMemory(dp.flat_layout(self.params)[1], dev, process_group)
            except RuntimeError as e:                      # raised on every rank together (no peer access / IPC refused)
                if not auto:
                    raise
                import sys
                print(f"scat_b200: {e}; gradient all-reduce falls back to NCCL", file=sys.stderr)
                self.comm = "nccl"
        # p.grad are views of one flat buffer
        self.bucket = dp.FlatGradBucket(self.params, flat=self.peer.flat if self.peer is not None else None)
        r = net.mask_rate
        self.n_masked = int(r * net.full_content) if (r >= 0.1 and r <= 0.9) else 0
        if x2_dtype not in ("fp32", "bf16"):
            raise ValueError(f"x2_dtype={x2_dtype!r}: 'fp32' or 'bf16'")
        # seam storage: x2 arrives, and x2.grad leaves, in the backbone's dtype
        self.x2_dtype = x2_dtype
        seam_t = torch.bfloat16 if x2_dtype == "bf16" else torch.float32
        self.cfg = net.config(self.n_masked, x2_dtype=x2_dtype)
        B = self.batch
        self.n_slots = int(input_slots)
        self.x2s = [torch.empty(B, 512, 28, 28, device=dev, dtype=seam_t) for _ in range(self.n_slots)]
        self.main_feats = [torch.empty(B, 1024, device=dev) for _ in range(self.n_slots)]
        # train.py:188-199 slices the label rows by their width: 105 = [63 3D | 42 2D], 166 = [61 pose | 63 3D | 42 2D]
        if label_width not in (105, 166):
            raise ValueError(f"label_width={label_width}: the reference's label rows are 105 or 166 wide (train.py:188-199)")
        self.label_width = int(label_width)
        self.labelss = [torch.empty(B, self.label_width, device=dev) for _ in range(self.n_slots)]
        self.copied: List[Optional[torch.cuda.Event]] = [None] * self.n_slots     # slot contents are ready
        self.consumed: List[Optional[torch.cuda.Event]] = [None] * self.n_slots   # slot may be overwritten
        self.pred = torch.empty(B, 66, device=dev)
        self.feat_visual = torch.empty(B, 21, 28, 28, device=dev)
        self.pl = torch.empty(B, 21, 28, 28, device=dev) if self.cfg.pl_reg else None
        self.losses = torch.zeros(4, device=dev)
        self.x2_grad = torch.empty(B, 512, 28, 28, device=dev, dtype=seam_t) if need_x2_grad else None
        self.main_feat_grad = torch.empty(B, 1024, device=dev) if need_main_feat_grad else None
        self.mask_dev = torch.zeros(max(self.n_masked, 1), dtype=torch.int32, device=dev)
        # the host-drawn indices travel through a ring of pinned slots, one asynchronous copy per step (a pageable source
        # would make copy_ synchronise the stream every step and expose the graph-launch latency)
        self._mask_host = torch.zeros(self._RING, max(self.n_masked, 1), dtype=torch.int32).pin_memory()
        self._mask_events: List[Optional[torch.cuda.Event]] = [None] * self._RING
        self._mask_turn = 0
        self._steps_since_poll = 0
        self.ws = SF.alloc_workspace(self.cfg, B, dev)
        self.last_mask = []
        self.use_graph = use_graph
        self.graphs = {}      # (slot, all-reduce in graph, optimiser in graph) -> CUDAGraph
        # phased issue (multi-rank): per slot three graphs, a communication stream, and the split points of the
        # bucket = first element of parameter 2 (transformer layer 0) and 13 (layer 1), in state_dict order
        self.phase_graphs: List[Optional[tuple]] = [None] * self.n_slots
        self.phased = bool(phased) and self.peer is None
        self.comm_stream = torch.cuda.Stream(device=dev) if self.phased else None
        self.split = self.bucket.offsets[13]
        self.split0 = self.bucket.offsets[2]
        self.split_ff = self.bucket.offsets[7]      # first element of layer 0's feed-forward half (its PreNorm weight)
        # peer-memory exchange hidden under the backward: the step is issued in its three phases inside ONE graph and the
        # part of the bucket each phase completes is summed on a second stream while the next phase computes
        self.overlap_exchange = bool(overlap_exchange) and self.peer is not None

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
    def _enqueue(self, slot: int = 0, phase: int = -1, ready=None):
        """Enqueue one fused step (or one phase of it) on the current stream (graph-capturable: no allocation, no
        sync).  ``ready(part, stream_ptr)``: the gradients-ready hook of scat_head_train_step_hooked (single call only)."""
        cfg = self.cfg
        d = cfg.desc(self.batch)
        pe = self.net.positionalEncoding.pe[0] if cfg.pos_embed else None
        labels = self.labelss[slot]
        if ready is not None:
            if phase != -1:
                raise ValueError("_enqueue: the gradients-ready hook belongs to the single-call step")
            failure = []

            def _cb(_user, part, stream):
                try:
                    ready(int(part), C.c_void_p(stream))
                    return 0
                except Exception as e:               # an exception must not unwind through the C frames
                    failure.append(e)
                    return 1
            cb = _lib.GRADS_READY_FN(_cb)
            rc = self.lib.scat_head_train_step_hooked(
                C.byref(d), ptr_array([p.data for p in self.params]), ptr(pe), ptr(self.net.mean_params.reshape(-1)),
                ptr(self.mask_dev) if self.n_masked else None, ptr(self.x2s[slot]), ptr(self.main_feats[slot]),
                ptr(labels), labels.shape[1], self.w3d, self.w2d, 1.0 / self.world, ptr(self.pred),
                ptr(self.feat_visual), ptr(self.pl), ptr(self.losses), ptr_array(self.bucket.views), ptr(self.x2_grad),
                ptr(self.main_feat_grad), ptr(self.ws), self.ws.numel(), C.cast(cb, C.c_void_p), None, SF.stream_ptr())
            if failure:
                raise failure[0]
            check(rc, "scat_head_train_step_hooked")
            return
        check(self.lib.scat_head_train_step_phase(
            C.byref(d), ptr_array([p.data for p in self.params]), ptr(pe), ptr(self.net.mean_params.reshape(-1)),
            ptr(self.mask_dev) if self.n_masked else None, ptr(self.x2s[slot]), ptr(self.main_feats[slot]),
            ptr(labels), labels.shape[1], self.w3d, self.w2d, 1.0 / self.world, ptr(self.pred),
            ptr(self.feat_visual), ptr(self.pl), ptr(self.losses), ptr_array(self.bucket.views), ptr(self.x2_grad),
            ptr(self.main_feat_grad), ptr(self.ws), self.ws.numel(), SF.stream_ptr(), phase), "scat_head_train_step")

    def _warm(self, slot: int):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):            # warm-up outside capture: function attributes, module loading
                self._enqueue(slot)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()

    def attach_optimizer(self, opt):
        """Make ``step(optimize=True)`` apply ``opt`` (a scat_b200.optim.HeadAdam over the same parameters) right after
        the gradient all-reduce -- inside the step's CUDA graph whenever the all-reduce is."""
        if opt.flat_params.data_ptr() != self.flat_params.data_ptr() or opt.n != self.flat_params.numel():
            raise ValueError("attach_optimizer: the optimiser does not own this step's parameters")
        self.opt = opt
        if self.peer is not None:          # a gradient exchange that gave up on a peer must not reach the weights
            opt.abort_flag = self.peer.error_word()
        self.graphs = {k: g for k, g in self.graphs.items() if not k[2]}

    def _enqueue_all(self, slot: int, ar: bool, optimize: bool):
        if ar and self.overlap_exchange:
            # The library calls back, while it enqueues the step, each time a part of the bucket is final: layers 1, 2 and
            # the regressor (bucket[split:], 6 MB) once layer 1's weight gradients are queued -- their sum runs under the
            # layer-0 backward; layer 0's feed-forward half (bucket[split_ff:split], 2.8 MB) in the middle of layer 0; its
            # attention half (bucket[split0:split_ff], 6.4 MB) -- summed under the conv passes; the mask token and the conv
            # weight (bucket[:split0], 46 KB: two cross-GPU flag barriers and little else) -- under the conv data gradient,
            # the last kernel of the step.  The exchanges are launches of the same kernel on the library's exchange
            # stream: same order on every rank, serialised among themselves, no phase joins on the main stream; all but the
            # last leave their closing cross-GPU barrier to the last.
            parts = ((self.split, None), (self.split_ff, self.split), (self.split0, self.split_ff), (0, self.split0))
            self._enqueue(slot, ready=lambda part, stream: self.peer.enqueue(stream, lo=parts[part][0], hi=parts[part][1],
                                                                               last=(part == 3)))
        else:
            self._enqueue(slot)
            if ar:
                self.peer.enqueue(SF.stream_ptr())
        if optimize:
            self.opt.enqueue(self.bucket.flat)

    def _capture(self, slot: int, ar: bool = False, optimize: bool = False):
        self._warm(slot)
        if ar:
            self.peer.enqueue(SF.stream_ptr())             # load the kernel outside capture (every rank does)
        if optimize:                                       # same for the optimiser kernel, without moving anything
            keep = [t.clone() for t in (self.flat_params, self.opt.exp_avg, self.opt.exp_avg_sq)]
            self.opt.enqueue(self.bucket.flat)
            for dst, src in zip((self.flat_params, self.opt.exp_avg, self.opt.exp_avg_sq), keep):
                dst.copy_(src)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue_all(slot, ar, optimize)
        self.graphs[(slot, ar, optimize)] = g
        return g

    def _capture_phases(self, slot: int):
        self._warm(slot)
        gs = []
        for phase in (0, 1, 2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(slot, phase)
            gs.append(g)
        self.phase_graphs[slot] = tuple(gs)

    def set_mask(self, mask_idx=None):
        """Draw (or take) the token mask on the host and stage it for the next step.  Consumes exactly one
        ``random.shuffle`` like the reference forward (hand_net.py:370-372)."""
        masked = self.net._draw_mask() if mask_idx is None else list(mask_idx)
        if len(masked) != self.n_masked:
            raise ValueError(f"mask has {len(masked)} indices, configuration expects {self.n_masked}")
        self.last_mask = masked
        if self.n_masked:
            k = self._mask_turn % self._RING
            self._mask_turn += 1
            if self._mask_events[k] is not None:
                self._mask_events[k].synchronize()       # the copy that read this pinned slot _RING steps ago has finished
            self._mask_host[k].copy_(torch.tensor(masked, dtype=torch.int32))
            self.mask_dev.copy_(self._mask_host[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._mask_events[k] = ev
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
            if labels.shape[1] != self.label_width:
                raise ValueError(f"load_inputs: labels are {labels.shape[1]} wide, this step was built for label_width="
                                 f"{self.label_width} (train.py:188-199 picks the 3D / 2D columns by the row width)")
            self.labelss[slot].copy_(labels, non_blocking=non_blocking)
            ev = torch.cuda.Event()
            ev.record(cur)
            self.copied[slot] = ev

    def step(self, allreduce: bool = True, slot: int = 0, optimize: bool = False):
        """Run fwd + pl VJP + loss + bwd on the inputs staged in ``slot`` (+ the gradient all-reduce, + the attached
        optimiser's update with ``optimize``); returns the device tensor losses[4] = [loss, l_3d, l_2d, l_pl] (local
        to this rank's shard)."""
        cur = torch.cuda.current_stream()
        if self.copied[slot] is not None:
            cur.wait_event(self.copied[slot])
        if optimize and self.opt is None:
            raise RuntimeError("step(optimize=True): attach_optimizer first")
        do_ar = allreduce and self.world > 1
        ar_in_graph = do_ar and self.peer is not None       # the all-reduce kernel is part of the step's graph
        opt_in_graph = optimize and (ar_in_graph or not do_ar)
        overlap = do_ar and self.phased
        if optimize:
            self.opt.advance()
        if overlap:
            # each phase is followed by the all-reduce, on the communication stream, of the part of the bucket it
            # finished, which overlaps the next phase
            if self.use_graph and self.phase_graphs[slot] is None:
                self._capture_phases(slot)
            parts = ((self.split, None), (self.split0, self.split), (0, self.split0))
            for phase in (0, 1, 2):
                if self.use_graph:
                    self.phase_graphs[slot][phase].replay()
                else:
                    self._enqueue(slot, phase)
                if phase < 2:
                    self.comm_stream.wait_stream(cur)
                    with torch.cuda.stream(self.comm_stream):
                        self.bucket.all_reduce(self.pg, lo=parts[phase][0], hi=parts[phase][1])
        elif self.use_graph:
            g = self.graphs.get((slot, ar_in_graph, opt_in_graph)) or self._capture(slot, ar_in_graph, opt_in_graph)
            g.replay()
        else:
            self._enqueue_all(slot, ar_in_graph, opt_in_graph)
        if self.n_slots > 1:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.consumed[slot] = ev
        if overlap:
            self.bucket.all_reduce(self.pg, hi=self.split0)
            cur.wait_stream(self.comm_stream)
        elif do_ar and not ar_in_graph:
            self.bucket.all_reduce(self.pg)
        if optimize and not opt_in_graph:
            self.opt.enqueue(self.bucket.flat)
        if self.peer is not None and do_ar:
            self._steps_since_poll += 1
            if self._steps_since_poll >= self.POLL_EVERY:
                self.check_peers()
        return self.losses

    def check_peers(self):
        """Raise if a gradient exchange gave up waiting for a peer (20 s bounded wait).  From that step on the exchange
        and the attached optimiser are no-ops on this rank, so the weights are those of the last complete step.
        Synchronises the device; step() calls it every POLL_EVERY steps."""
        self._steps_since_poll = 0
        if self.peer is not None and self.peer.timed_out():
            raise RuntimeError("scat_b200: a rank never arrived at the gradient all-reduce (20 s bounded wait); the "
                               "exchange and the optimiser update were skipped from that step on")


    def close(self):
        """Release the peer-mapped gradient bucket (multi-rank runs; every rank must call it).  The parameters'
        ``.grad`` views die with it, so this is the last call on the object."""
        if self.peer is not None:
            self.graphs, self.phase_graphs = {}, [None] * self.n_slots
            for p in self.params:
                p.grad = None
            self.bucket = None
            self.peer.close()
            self.peer = None


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
