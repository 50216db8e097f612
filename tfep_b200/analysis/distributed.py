"""Batch-sharded (T)FEP estimator and bootstrap: one process per GPU, contiguous shards of the work values.

The reference has no multi-process path; this is the one parallelism the build adds.
Every sample is independent through the flow, so nothing is exchanged until the estimator:

* estimator: each rank reduces its shard to ``(max, sum exp(v - max))`` with the single-pass kernel; ONE
  all-gather of 2 doubles per rank and a rescale combine them (``combine_lse_partials``);
* bootstrap: every rank walks the SAME global index stream (the reference's MT19937 stream, or Philox) and
  sums the draws that fall into its shard against its own exp table; the per-resample sums are rescaled to the
  global maximum and added with ONE all-reduce of ``n_resamples`` doubles (``combine_bootstrap_sums``).

The collective helpers work on CPU tensors with the gloo backend too (that is how the host logic is tested
without GPUs); the local reductions are CUDA kernels.
"""

import torch
import torch.distributed as dist

from .. import _ops
from .bootstrap import _fused_kT, _generator_to_state, _state_to_generator
from .estimator import _log_n, combine_partials


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def combine_lse_partials(partial, group=None):
    """All-gather the ``(max, sum)`` pair of every rank and combine: returns ``(max, sum)`` of the union."""
    if _world(group) == 1:
        return partial[0], partial[1]
    parts = [torch.empty_like(partial) for _ in range(_world(group))]
    dist.all_gather(parts, partial.contiguous(), group=group)
    return combine_partials(torch.stack(parts))


def combine_bootstrap_sums(sums, local_max, group=None):
    """``sums[r] = sum over this rank's shard of exp(v - local_max)`` -> global ``(sums, max)``."""
    if _world(group) == 1:
        return sums, local_max
    gmax = local_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    scaled = sums * torch.exp(local_max - gmax)
    dist.all_reduce(scaled, op=dist.ReduceOp.SUM, group=group)
    return scaled, gmax


def _total(n_local, device, group):
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    if _world(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def fep_estimator_sharded(work_shard, kT=1.0, group=None):
    """``-kT logsumexp(-w / kT - log n)`` over the union of all ranks' shards (same value on every rank)."""
    m, s = combine_lse_partials(_ops.lse(work_shard, -1.0 / kT), group)
    n = _total(work_shard.numel(), work_shard.device, group)
    return (-kT * (m + torch.log(s) - _log_n(n))).to(work_shard.dtype)


def bootstrap_statistics_sharded(work_shard, shard_offset, n_total, kT=1.0, n_resamples=9999, batch=None, generator=None,
                                 rng='mt19937', group=None):
    """Per-resample FEP estimates of the global data set; every rank returns the same ``(n_resamples,)`` tensor.

    ``work_shard`` is the contiguous slice ``[shard_offset, shard_offset + len)`` of the global work values.
    All ranks must pass generators in the same state (same seed): they draw the same global indices.
    """
    dev = work_shard.device
    scale = -1.0 / kT
    o = _ops.lse(work_shard, scale)
    e = _ops.exp_table(work_shard, scale, o[:1])
    sums = torch.empty(n_resamples, dtype=torch.float64, device=dev)
    if rng == 'philox':
        seed = int(torch.randint(0, 2**62, (1,), generator=generator).item())
        for k in range(0, n_resamples, 65535):
            nb = min(65535, n_resamples - k)
            sums[k:k + nb] = _ops.bootstrap_sums(e, n_total, nb, n_total, None, seed, k * ((n_total + 3) // 4),
                                                 shard_lo=shard_offset)
    else:
        gen = torch.default_generator if generator is None else generator
        state = _generator_to_state(gen).to(dev)
        batch = max(1, min(n_resamples if batch is None else batch, (1 << 28) // max(n_total, 1) or 1))
        idx = torch.empty(batch * n_total, dtype=torch.int32, device=dev)
        for k in range(0, n_resamples, batch):
            nb = min(batch, n_resamples - k)
            _ops.mt19937_indices(state, nb * n_total, n_total, out=idx)
            sums[k:k + nb] = _ops.bootstrap_sums(e, n_total, nb, n_total, idx[:nb * n_total].view(nb, n_total),
                                                 shard_lo=shard_offset)
        _state_to_generator(gen, state)
    sums, gmax = combine_bootstrap_sums(sums, o[0], group)
    return (-kT * (gmax + torch.log(sums) - _log_n(n_total))).to(work_shard.dtype)
