"""Batch-sharded training across GPUs: one process per GPU, weights replicated, contiguous batch shards.

The flow has no cross-sample coupling, so the only exchange of a training step is the sum of the parameter
gradients: every rank back-propagates ``loss_local * n_local / n_global`` (or the mean
over equal shards) and the gradients are added with ONE all-reduce of a flat bucket -- NCCL over NVLink on the
GPU box, gloo in the CPU tests.  The reference has no multi-process path; this is the one parallelism the build adds.
"""

import torch
import torch.distributed as dist


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers (one flat broadcast per dtype)."""
    if _world(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for ts in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def allreduce_gradients(module, average=True, group=None):
    """Sum (or average) the gradients of ``module`` over all ranks with one all-reduce of a flat bucket.

    Parameters without a gradient on this rank contribute zeros (every rank must own the same parameters).
    Returns the number of elements reduced."""
    world = _world(group)
    params = [p for p in module.parameters() if p.requires_grad]
    if world == 1 or not params:
        return 0
    by_dtype = {}
    for p in params:
        by_dtype.setdefault((p.dtype, p.device), []).append(p)
    total = 0
    for ps in by_dtype.values():
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for p in ps:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        total += flat.numel()
    return total


def shard_bounds(n, rank=None, world=None, group=None):
    """Contiguous shard ``[lo, hi)`` of ``n`` samples owned by ``rank``."""
    world = _world(group) if world is None else world
    rank = (dist.get_rank(group) if world > 1 else 0) if rank is None else rank
    return rank * n // world, (rank + 1) * n // world
