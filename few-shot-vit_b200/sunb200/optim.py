"""Fused multi-tensor optimizers and the learning-rate schedules of the reference's training scripts.

FusedSGD / FusedAdamW are torch.optim.Optimizer subclasses (param_groups, state_dict, zero_grad and torch's LR schedulers
keep working) whose step() is ONE native launch over every parameter tensor (csrc/optim.cu):
  * FusedSGD   = torch.optim.SGD(lr, momentum, weight_decay) as utils.make_optimizer builds it
                 (reference meta_tuning_sun_m/utils/__init__.py:128-139, train_meta_warmup.py:140)
  * FusedAdamW = AdamW(betas, eps, decoupled weight decay) (reference sun_meta_training/offline.py:229)
The tensor table is passed to the kernel by value (no copy at step time); the hyper-parameters and the AdamW step counter
live in a small device tensor, so a CUDA graph that captured step() stays valid when the scheduler changes the learning rate.

CosineLRScheduler / MultiStepLRScheduler restate the timm schedulers the reference imports (offline.py:231,
train_meta_warmup.py:141).  timm is not installed in this image: parity with timm is unpinned, the formulas are timm's
published ones (linear warm-up from warmup_lr_init over warmup_t epochs; cosine to lr_min over t_initial with
cycle_limit 1; step decay decay_rate ** bisect_right(decay_t, t + 1)).
"""
from __future__ import annotations

import bisect
import ctypes as C
import math
from typing import List

import torch

from . import native as N


class _FusedBase(torch.optim.Optimizer):
    _adamw = False

    def _init_tables(self):
        self._hp = None
        self._hp_host = None

    @staticmethod
    def _table(params, grads, bufs1, bufs2):
        """Host-side tensor table; the native launcher passes it to the kernel by value (nothing is copied at step time)."""
        arr = (N.OptTensor * len(params))()
        for i, (p, g, m, v) in enumerate(zip(params, grads, bufs1, bufs2)):
            arr[i].p, arr[i].g, arr[i].m = p.data_ptr(), g.data_ptr(), m.data_ptr()
            arr[i].v = None if v is None else v.data_ptr()
            arr[i].n = p.numel()
        return arr

    def _hyper(self, dev, values):
        """Device hyper-parameter block; re-uploaded only when a value (e.g. the learning rate) changed.  Slot 5 (the
        AdamW step counter) is owned by the device."""
        if self._hp is None:
            self._hp = torch.zeros(8, dtype=torch.float32, device=dev)
            self._hp_host = [None] * 5
        if self._hp_host != values:
            self._hp_stage = torch.tensor(values, dtype=torch.float32).pin_memory()
            self._hp[:5].copy_(self._hp_stage, non_blocking=True)
            self._hp_host = list(values)
        return self._hp


class FusedSGD(_FusedBase):
    def __init__(self, params, lr, momentum=0.9, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedSGD supports a single parameter group")
        self._init_tables()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        grp = self.param_groups[0]
        params = [p for p in grp["params"] if p.grad is not None]
        if not params:
            return loss
        N.require_cuda(*params)
        grads, bufs = [], []
        for p in params:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("FusedSGD needs contiguous fp32 parameters and gradients")
            st = self.state[p]
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(p)      # zero start == torch's "first step: buf = grad"
            grads.append(p.grad)
            bufs.append(st["momentum_buffer"])
        arr = self._table(params, grads, bufs, [None] * len(params))
        hp = self._hyper(params[0].device, [float(grp["lr"]), float(grp["momentum"]), float(grp["weight_decay"]), 0.0, 0.0])
        N.check(N.lib().sunb_fused_sgd(arr, len(params), hp.data_ptr(), N.current_stream()), "sunb_fused_sgd")
        return loss


class FusedAdamW(_FusedBase):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group")
        self._init_tables()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        grp = self.param_groups[0]
        params = [p for p in grp["params"] if p.grad is not None]
        if not params:
            return loss
        N.require_cuda(*params)
        grads, m1, m2 = [], [], []
        for p in params:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
            grads.append(p.grad)
            m1.append(st["exp_avg"])
            m2.append(st["exp_avg_sq"])
        arr = self._table(params, grads, m1, m2)
        b1, b2 = grp["betas"]
        hp = self._hyper(params[0].device, [float(grp["lr"]), float(b1), float(grp["weight_decay"]), float(b2), float(grp["eps"])])
        N.check(N.lib().sunb_fused_adamw(arr, len(params), hp.data_ptr(), N.current_stream()), "sunb_fused_adamw")
        return loss


class _EpochScheduler:
    """timm-style scheduler: `step(epoch)` sets the learning rate of every param group for that epoch."""

    def __init__(self, optimizer, warmup_t=0, warmup_lr_init=0.0):
        self.optimizer = optimizer
        self.base_values = [g["lr"] for g in optimizer.param_groups]
        self.warmup_t, self.warmup_lr_init = warmup_t, warmup_lr_init
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self._set([warmup_lr_init for _ in self.base_values])

    def _set(self, values: List[float]):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def _lrs(self, t):
        raise NotImplementedError

    def get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        return self._lrs(t)

    def step(self, epoch):
        self._set(self.get_lr(epoch))


class CosineLRScheduler(_EpochScheduler):
    """timm.scheduler.CosineLRScheduler as offline.py:231 configures it (t_initial = max_epoch, cycle_decay 0.1,
    cycle_limit 1, warmup_prefix False, lr_min 0)."""

    def __init__(self, optimizer, t_initial, lr_min=0.0, cycle_decay=1.0, cycle_limit=1, warmup_t=0, warmup_lr_init=0.0):
        super().__init__(optimizer, warmup_t, warmup_lr_init)
        self.t_initial, self.lr_min, self.cycle_decay, self.cycle_limit = t_initial, lr_min, cycle_decay, cycle_limit

    def _lrs(self, t):
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        gamma = self.cycle_decay ** i
        if i < self.cycle_limit:
            return [self.lr_min + 0.5 * (v * gamma - self.lr_min) * (1 + math.cos(math.pi * t_curr / self.t_initial))
                    for v in self.base_values]
        return [self.lr_min for _ in self.base_values]


class MultiStepLRScheduler(_EpochScheduler):
    """timm.scheduler.MultiStepLRScheduler as train_meta_warmup.py:141 configures it (decay_rate 0.5, warmup_lr_init 1e-5,
    warmup_t 3)."""

    def __init__(self, optimizer, decay_t, decay_rate=1.0, warmup_t=0, warmup_lr_init=0.0):
        super().__init__(optimizer, warmup_t, warmup_lr_init)
        self.decay_t, self.decay_rate = list(decay_t), decay_rate

    def _lrs(self, t):
        k = bisect.bisect_right(self.decay_t, t + 1)
        return [v * (self.decay_rate ** k) for v in self.base_values]
