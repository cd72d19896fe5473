"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the collectives.

Evaluation: episodes are independent -> `shard_range` gives rank r the episodes [r*E/N, (r+1)*E/N); no data-path
collective, one final gather of per-episode results (reference: test_phase/test_few_shot.py:79-94 has no cross-episode
state).  Meta-tuning: the episode axis of each batch is sharded exactly as nn.DataParallel scatters it
(meta_tuning_sun_m/train_meta.py:128-129,168); BatchNorm statistics stay per replica (un-synchronised, as under
DataParallel); the only collective is one all-reduce(mean) of the 12,531,393 fp32 gradients per step.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_episodes(x_shot: torch.Tensor, x_query: torch.Tensor, rank: int, world: int):
    """Slice the leading episode axis of (x_shot [E,...], x_query [E,...]) for this rank."""
    lo, hi = shard_range(x_shot.shape[0], rank, world)
    return x_shot[lo:hi], x_query[lo:hi]


class GradAllReducer:
    """One flat fp32 bucket for all gradients: copy-in, a single NCCL all-reduce (NVLS on NVSwitch), copy-out as the mean.
    The bucket is allocated once; `params` order fixes the layout on every rank."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Make every .grad a view of the bucket so backward writes land in place (no copy-in)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self, group=None):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        for p, v in zip(self.params, self.views):          # gradients that autograd re-allocated are copied back in
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None):
    """Rank-0 parameters and BatchNorm buffers are authoritative (DataParallel keeps replica 0's buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def sync_bn_buffers(module: torch.nn.Module, src: int = 0, group=None):
    """Re-broadcast rank `src`'s BatchNorm buffers (running_mean / running_var / num_batches_tracked).  During multi-process
    meta-tuning every rank updates its own running statistics (un-synchronised BN, as each DataParallel replica does);
    DataParallel then keeps replica 0's.  Call this at epoch end, before a rank-sharded evaluation and before saving a
    checkpoint, so that every rank evaluates / stores rank 0's statistics."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for b in module.buffers():
        dist.broadcast(b.data, src=src, group=group)
    for m in module.modules():
        if hasattr(m, "invalidate_packed_weights"):
            m.invalidate_packed_weights()


class GradComm:
    """Overlapped data-parallel gradient reduction for the native backward pass (sunb200/train.py): every ready slice of
    the flat gradient buffer is all-reduced (sum) on a dedicated stream while the compute stream keeps running the
    remaining backward kernels; `finish` joins the streams and turns the sums into means."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def all_reduce_async(self, flat_slice: torch.Tensor):
        if self.world == 1:
            return
        self.stream.wait_stream(torch.cuda.current_stream())       # the slice's producers have been enqueued
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self, flat: torch.Tensor):
        if self.world == 1:
            return
        torch.cuda.current_stream().wait_stream(self.stream)
        flat.div_(self.world)
