"""Fused Adam for the head: ``optim.Adam(self.net.parameters(), lr=self.lr)`` +
``optimizer.step()`` of the reference (train.py:60,209) as ONE kernel over flat parameter / gradient / moment buffers
(csrc/adam.cu, ``scat_adam_step``), and the reference's learning-rate warm-up (train.py:61-63,134).

``HeadAdam`` is a ``torch.optim.Optimizer``: ``param_groups[0]["lr"]`` is what torch's schedulers drive,
``state_dict()`` has torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so checkpoints
move between the two.  It covers the head's tensors (``EncoderTransformer.head_parameters()``); the cuDNN backbone
keeps a PyTorch optimiser (north_star: the backbone stays a PyTorch feature producer).

Used with ``HeadTrainStep.attach_optimizer`` the update is captured into the same CUDA graph as forward, backward and
the gradient all-reduce; step count and learning rate then live in device memory and are refreshed by one small
asynchronous host-to-device copy per step (pinned staging ring, no stream synchronisation).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, dp
from ._lib import check, ptr, stream_ptr


def gradual_warmup_lr(base_lr: float, epoch: int, total_epoch: int = 15) -> float:
    """Learning rate the reference uses during 0-based ``epoch``: GradualWarmupScheduler(multiplier=1,
    total_epoch=15, after_scheduler=StepLR(step_size=10, gamma=1)) stepped with ``epoch + 1`` (train.py:61-63,134) --
    linear from base_lr/15 to base_lr over the first 15 epochs, constant afterwards."""
    last_epoch = epoch + 1
    if last_epoch > total_epoch:
        return base_lr
    return base_lr * (float(last_epoch) / total_epoch)


class WarmupSchedule:
    """``scheduler_warmup.step(epoch + 1)`` of the reference loop for any optimiser with ``param_groups``."""

    def __init__(self, optimizer, total_epoch: int = 15):
        self.optimizer, self.total_epoch = optimizer, int(total_epoch)
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.step(0)                                        # _LRScheduler.__init__ steps once: rate 0 until step(1)

    def step(self, last_epoch: int):
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = gradual_warmup_lr(base, last_epoch - 1, self.total_epoch) if last_epoch > 0 else 0.0


class HeadAdam(torch.optim.Optimizer):
    _RING = 8          # pinned staging slots for (step, learning rate): the host may run this many steps ahead

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError(f"HeadAdam: lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("HeadAdam: one parameter group (the head's tensors)")
        ps = self.param_groups[0]["params"]
        if ps[0].device.type != "cuda" or any(p.dtype != torch.float32 for p in ps):
            raise RuntimeError("HeadAdam: fp32 parameters on a CUDA device (no CPU path)")
        self.lib = _lib.load()
        self.flat_params = dp.flatten_parameters(ps)         # p.data become views of one buffer
        self.n = self.flat_params.numel()
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.step_count = 0
        dev = self.flat_params.device
        self._sched_dev = torch.zeros(2, dtype=torch.int32, device=dev)      # [step, bits of the fp32 learning rate]
        self._step_dev = self._sched_dev[0:1]
        self._lr_dev = self._sched_dev[1:2].view(torch.float32)
        # (step, learning-rate bits) travel through a small ring of pinned host words, one asynchronous copy per step:
        # a pageable source would make copy_ synchronise the stream, i.e. drain the device every step
        self._sched_host = torch.zeros(self._RING, 2, dtype=torch.int32).pin_memory() if dev.type == "cuda" else None
        self._sched_events = [None] * self._RING
        self._sched_turn = 0
        # device address of a word that, when non-zero, turns the update into a no-op (HeadTrainStep wires it to the
        # gradient exchange's time-out flag)
        self.abort_flag = None
        self._link_state()

    def _link_state(self):
        ps = self.param_groups[0]["params"]
        for p, off in zip(ps, dp.flat_layout(ps)[0]):
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.exp_avg[off: off + n].view(p.shape),
                             "exp_avg_sq": self.exp_avg_sq[off: off + n].view(p.shape)}

    def _grads(self) -> torch.Tensor:
        g = dp.flat_gradients(self.param_groups[0]["params"])
        if g is None:
            raise RuntimeError("HeadAdam: the gradients must be the slices of one flat bucket (dp.FlatGradBucket / "
                               "HeadTrainStep) -- was a .grad replaced or set to None?")
        return g

    def _launch(self, grads, lr, step, lr_dev, step_dev):
        g = self.param_groups[0]
        check(self.lib.scat_adam_step(ptr(self.flat_params), ptr(grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n,
                                      lr, g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], step, lr_dev,
                                      step_dev, self.abort_flag, stream_ptr()), "scat_adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        """optimizer.step(): one kernel on the current stream, count and rate passed by value."""
        loss = closure() if closure is not None else None
        self.step_count += 1
        self._launch(self._grads(), float(self.param_groups[0]["lr"]), self.step_count, None, None)
        return loss

    # ---- graph-replayable form: advance() on the host each step, enqueue() captured once ----
    def advance(self):
        """Count one step and stage (learning rate, step) for the device on the current stream."""
        self.step_count += 1
        k = self._sched_turn % self._RING
        self._sched_turn += 1
        if self._sched_events[k] is not None:
            self._sched_events[k].synchronize()          # the copy that last read this pinned slot (RING steps ago) is done
        slot = self._sched_host[k]
        slot[0] = self.step_count
        slot[1:2].view(torch.float32)[0] = float(self.param_groups[0]["lr"])
        self._sched_dev.copy_(slot, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._sched_events[k] = ev

    def enqueue(self, grads: torch.Tensor = None):
        """The update with count / rate read from device memory (capturable; call advance() before each replay)."""
        self._launch(self._grads() if grads is None else grads, 0.0, 0, ptr(self._lr_dev), ptr(self._step_dev))

    def zero_grad(self, set_to_none: bool = True):
        """The fused step overwrites the gradient bucket at the start of every backward, so there is nothing to
        release; ``set_to_none`` must not detach the bucket views and is treated as a plain zero-fill."""
        g = dp.flat_gradients(self.param_groups[0]["params"])
        if g is not None:
            g.zero_()

    def state_dict(self):
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.step_count))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                  # torch.optim.Adam layout -> per-parameter copies
        steps = set()
        ps = self.param_groups[0]["params"]
        for p, off in zip(ps, dp.flat_layout(ps)[0]):
            n = p.numel()
            st = self.state.get(p, None)
            if st:
                self.exp_avg[off: off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off: off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"HeadAdam: parameters disagree on the step count {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        self._link_state()
