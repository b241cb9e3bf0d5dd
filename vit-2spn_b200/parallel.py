"""Data-parallel plumbing for the SSP step: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

The reference's DDP branch is dead code (``use_distributed = False``, ref:ssp_vit2spn_tiny.py:21-25,
169-172), so the only behaviour to match is the mathematical one: N ranks x B samples
must equal one process with the concatenated N*B batch.  The reference loss is a plain mean of
per-sample cosines (no negatives), hence the path shards with ONE collective per
optimizer step: the average of the gradients.  Gradients live in three flat fp32 buffers (online
backbone 1, online backbone 2, heads), so the all-reduce is three large NCCL calls on flat memory —
no per-tensor bucketing pass.  EMA and Adam are replicated (deterministic given equal gradients).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def gradient_buckets(model):
    """The flat gradient ranges of a ``DualStreamNetwork`` in reverse execution order (heads first,
    as their gradients are final first): views, no copies."""
    stores = model._stores()[:2]
    buckets = [model._head_store.flat_grad]
    for s in stores[::-1]:
        buckets.append(s.flat_grad[:s.active_numel] if s.flat_grad is not None else None)
    if any(b is None for b in buckets):
        raise RuntimeError("gradient_buckets: run a backward pass first (flat gradient buffers are lazy)")
    return buckets


def allreduce_buckets(buckets, group=None, average=True, async_op=True):
    """Sum (or average) every bucket over the process group.  Returns after all collectives are
    enqueued and waited on the current stream (NCCL) / completed (gloo)."""
    if not dist.is_available() or not dist.is_initialized():
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=async_op) for b in buckets]
    if async_op:
        for w in works:
            w.wait()
    if average:
        for b in buckets:
            b.mul_(1.0 / world)
    return world


def allreduce_gradients(model, group=None, optimizer=None):
    """Gradient all-reduce for one optimizer step.  With a ``FusedAdam`` the 1/world factor is folded
    into the Adam kernel (``optimizer.grad_multiplier``) instead of a separate pass over the gradients."""
    buckets = gradient_buckets(model)
    fold = optimizer is not None and hasattr(optimizer, "grad_multiplier")
    world = allreduce_buckets(buckets, group=group, average=not fold)
    if fold:
        optimizer.grad_multiplier = 1.0 / world
    return world


class OverlappedGradSync:
    """Bucketed gradient all-reduce overlapped with the backward pass (north_star; the reference's dead
    DDP branch ref:ssp_vit2spn_tiny.py:21-25,169-172 would have done the same through DDP's buckets).

    Buckets follow the flat gradient layout in reverse execution order: the heads, then block ranges
    ``[hi, lo)`` of both online backbones (the last range carries the embeddings).  ``model.ssp_step(...,
    grad_sync=self)`` issues the backward pass range by range and calls ``range_ready``: the all-reduce of a range is
    enqueued behind the kernels that produced it (c10d orders a collective after the work already queued on the
    current stream) and runs on NCCL's stream while the next range computes.  ``finish`` joins before the optimizer
    step and folds 1/world into ``FusedAdam.grad_multiplier``.

    The compute kernels are persistent grids of one 227-KB-shared-memory CTA per SM, so a NCCL kernel cannot co-reside
    with them: while collectives are in flight the library sizes its grids to ``#SMs - comm_sms``
    (``v2s_set_sm_limit``) and NCCL (``NCCL_MAX_CTAS``, set by the caller before init) runs on the SMs left free.
    Shrinking the grids costs ``comm_sms / #SMs`` of every kernel launched meanwhile, so the limit is set as late as
    possible: at the first ``range_ready``, together with the (small) head bucket.  Measured on 8 B200s (ms/step, 6.96
    on one GPU): 8 SMs for NCCL 7.19-7.21 with ranges (12,8,4,0) or (12,6,2,0); 4 SMs 7.25; 2 SMs 7.75 (NCCL too slow
    to hide); four ranges (12,7,3,1,0) with 4 SMs 7.43."""

    def __init__(self, model, group=None, splits=(8, 4), comm_sms=8):
        from . import _lib
        self.model, self.group, self.comm_sms = model, group, int(comm_sms)
        edges = [12] + [int(s) for s in splits] + [0]
        if any(a <= b for a, b in zip(edges, edges[1:])):
            raise ValueError("splits must be strictly decreasing block indices in (0, 12)")
        self.ranges = list(zip(edges[:-1], edges[1:]))
        offs = _lib.backbone_layout()
        self._layer0, self._layer_numel = offs[4], offs[20] - offs[4]
        self._works = []
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._sms = None

    def _slice(self, store, hi, lo):
        a = 0 if lo == 0 else self._layer0 + lo * self._layer_numel
        b = self._layer0 + hi * self._layer_numel
        return store.flat_grad[a:b]

    def begin(self):
        self._works = []
        self._limited = False
        self._heads_pending = False

    def _limit(self):
        from . import _lib
        if not self._limited and self.world > 1 and self.comm_sms > 0:
            if self._sms is None:
                self._sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            _lib.check(_lib.lib.v2s_set_sm_limit(self._sms - self.comm_sms), "set_sm_limit")
        self._limited = True

    def _reduce(self, t):
        if self.world > 1:
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def heads_ready(self):
        self._heads_pending = True      # sent with the first block range (see the class note)

    def range_ready(self, hi, lo):
        self._limit()
        if self._heads_pending:
            self._reduce(self.model._head_store.flat_grad)
            self._heads_pending = False
        for s in self.model._stores()[:2][::-1]:
            self._reduce(self._slice(s, hi, lo))

    def finish(self, optimizer=None):
        """Join every collective on the current stream; the gradients hold the SUM over ranks (1/world is folded into
        the optimizer if given, else applied here)."""
        from . import _lib
        if getattr(self, "_heads_pending", False):
            self._reduce(self.model._head_store.flat_grad)
            self._heads_pending = False
        for w in self._works:
            w.wait()
        self._works = []
        _lib.check(_lib.lib.v2s_set_sm_limit(0), "set_sm_limit")
        if self.world > 1:
            if optimizer is not None and hasattr(optimizer, "grad_multiplier"):
                optimizer.grad_multiplier = 1.0 / self.world
            else:
                for b in gradient_buckets(self.model):
                    b.mul_(1.0 / self.world)
        return self.world


def broadcast_parameters(model, src=0, group=None):
    """Identical replicas at start (the reference relies on a fixed seed, ref:ssp_vit2spn_tiny.py:47-50)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for s in model._stores() + [model._head_store]:
        s.ensure()
        dist.broadcast(s.flat, src, group=group)
        s.lp_fresh = False


def shard_batch(x, rank, world):
    """Rank r owns samples [r*B/world, (r+1)*B/world) of a global batch."""
    n = x.shape[0]
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return x[rank * per:(rank + 1) * per]
