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
