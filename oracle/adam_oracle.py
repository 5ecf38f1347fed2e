"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's optimiser step.

The reference trains with ``optim.Adam(self.net.parameters(), lr=self.lr)`` (/root/reference/train.py:60, default
betas (0.9, 0.999), eps 1e-8, weight_decay 0, no amsgrad; lr 1e-4, config.py:34), steps it once per batch
(train.py:209) and scales the learning rate once per epoch with
``GradualWarmupScheduler(optimizer, multiplier=1, total_epoch=15, after_scheduler=StepLR(step_size=10, gamma=1))``
stepped as ``scheduler_warmup.step(epoch + 1)`` (train.py:61-63,134).

* ``adam_step`` restates torch.optim.Adam's single-tensor path (torch/optim/adam.py ``_single_tensor_adam``, the
  non-capturable branch: scalar bias corrections in Python floats, ``exp_avg.lerp_``, ``exp_avg_sq.mul_().addcmul_``,
  ``param.addcdiv_``) in numpy float32.  Parity pin: oracle/make_golden.py runs torch.optim.Adam itself on the CPU and
  stores tests/golden/adam.npz; tests/test_oracle_golden.py re-checks this file against it.
* ``gradual_warmup_lr`` restates the published algorithm of the ``warmup_scheduler`` package
  (ildoonet/pytorch-gradual-warmup-lr; requirements.txt:126 pulls it from a git URL with no pinned revision, and it
  is not installed here): for ``multiplier == 1`` the rate is ``base_lr * last_epoch / total_epoch`` while
  ``last_epoch <= total_epoch`` and the after-scheduler's rate (StepLR with gamma 1: ``base_lr``) afterwards.
  PARITY UNPINNED for this function: the package is absent, only the reference's call sites anchor it.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import numpy as np


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One Adam update of float32 arrays, in place; ``step`` is the 1-based count AFTER the increment."""
    f, d = np.float32, np.float64

    def fma(a, b, c):                                         # float32 fused multiply-add (exact product in float64)
        return (a.astype(d) * b.astype(d) + c.astype(d)).astype(f)

    if weight_decay != 0:
        g = fma(np.full_like(p, weight_decay), p, g)          # grad.add(param, alpha=weight_decay)
    # The rounding points below are the ones ATen's vectorised CPU kernels were observed to use (fused multiply-adds
    # in lerp and addcmul, value * t1 before the division in addcdiv): bit-identical to torch.optim.Adam on all but
    # ~0.01% of elements, which differ by one ulp.
    m[...] = fma(np.full_like(m, 1 - beta1), g - m, m)        # exp_avg.lerp_(grad, 1 - beta1), weight < 0.5 branch
    v *= f(beta2)                                             # exp_avg_sq.mul_(beta2)
    v[...] = fma(f(1 - beta2) * g, g, v)                      #            .addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p += (f(-step_size) * m) / denom                          # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v


def gradual_warmup_lr(base_lr: float, epoch: int, total_epoch: int = 15) -> float:
    """Learning rate in force during 0-based ``epoch`` of the reference loop (train.py:134 calls step(epoch + 1))."""
    last_epoch = epoch + 1
    if last_epoch > total_epoch:
        return base_lr                                        # StepLR(step_size=10, gamma=1) keeps base_lr
    return base_lr * (float(last_epoch) / total_epoch)
