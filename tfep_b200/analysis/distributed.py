"""Batch-sharded (T)FEP estimator and bootstrap: one process per GPU, contiguous shards of the work values.

The reference has no multi-process path; this is the one parallelism the build adds.
Every sample is independent through the flow, so nothing is exchanged until the estimator:

* estimator: each rank reduces its shard to ``(max, sum exp(v - max))`` with the single-pass kernel; ONE
  all-gather of 2 doubles per rank and a rescale combine them (``combine_lse_partials``);
* bootstrap, reference index stream (MT19937): every rank walks the SAME global stream and sums the draws that fall
  into its shard against its own exp table;
* bootstrap, Philox: the shards are cut into L2-sized cells, every rank draws the same Multinomial counts of
  draws per cell (same seed) and generates only the draws of its own cells (``stratified_counts``) -- no draw is
  generated twice, so the work divides by the number of ranks;
* either way the per-resample partial sums are combined in the log domain with a MAX and a SUM all-reduce of
  ``n_resamples`` doubles (``combine_bootstrap_sums``), exact for data of any dynamic range.

The collective helpers work on CPU tensors with the gloo backend too (that is how the host logic is tested
without GPUs); the local reductions are CUDA kernels.
"""

import torch
import torch.distributed as dist

from .. import _ops
from .bootstrap import (_fused_kT, _generator_to_state, _state_to_generator, philox_cell_sums, repair_underflow,
                        stratified_counts, table_cells)
from .estimator import _log_n, combine_partials


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def combine_lse_partials(partial, group=None):
    """All-gather the ``(max, sum)`` pair of every rank and combine: returns ``(max, sum)`` of the union."""
    world = _world(group)
    if world == 1:
        return partial[0], partial[1]
    # one collective into one tensor (the call is latency bound: 16 bytes per rank)
    parts = torch.empty((world, 2), dtype=partial.dtype, device=partial.device)
    if partial.is_cuda:
        dist.all_gather_into_tensor(parts, partial.contiguous(), group=group)
    else:                                  # gloo (CPU tests) has no all_gather_into_tensor
        chunks = list(parts.unbind(0))
        dist.all_gather(chunks, partial.contiguous(), group=group)
        parts = torch.stack(chunks)
    return combine_partials(parts)


def combine_bootstrap_sums(sums, local_ref, group=None):
    """``sums[r] = sum over this rank's shard of exp(v - local_ref[r])`` (``local_ref``: a scalar or one reference per
    resample) -> global ``(sums, ref)`` with one reference per resample.  Combined in the log domain -- a MAX and a SUM
    all-reduce of ``n_resamples`` doubles -- so that a rank whose shard lies far below another's neither underflows nor
    drags the others down (the reference takes a per-row logsumexp, bootstrap.py:227-231)."""
    if _world(group) == 1:
        return sums, local_ref
    log_local = torch.where(sums > 0, local_ref + torch.log(sums.clamp_min(1e-300)), torch.full_like(sums, -float('inf')))
    ref = log_local.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.MAX, group=group)
    scaled = torch.where(torch.isfinite(log_local), torch.exp(log_local - ref), torch.zeros_like(sums))
    dist.all_reduce(scaled, op=dist.ReduceOp.SUM, group=group)
    return scaled, ref


def _total(n_local, device, group):
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    if _world(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def shard_cells(shard_offset, shard_len, n_total, device, group=None):
    """The L2-sized cells of every rank's shard, in global index order, and the rank owning each: ``(cells, owner)``.
    One all-gather of two integers per rank."""
    mine = torch.tensor([shard_offset, shard_len], dtype=torch.int64, device=device)
    if _world(group) > 1:
        parts = [torch.empty_like(mine) for _ in range(_world(group))]
        dist.all_gather(parts, mine, group=group)
        shards = [(int(p[0]), int(p[1]), r) for r, p in enumerate(torch.stack(parts).cpu())]
    else:
        shards = [(int(shard_offset), int(shard_len), 0)]
    cells, owner, pos = [], [], 0
    for off, ln, r in sorted(shards):
        if off != pos:
            raise ValueError('the shards of the ranks must tile [0, n_total) without gaps or overlaps')
        cs = table_cells(off, off + ln) if ln > 0 else []
        cells += cs
        owner += [r] * len(cs)
        pos = off + ln
    if pos != n_total:
        raise ValueError('the shards of the ranks must tile [0, n_total) without gaps or overlaps')
    return cells, owner


def fep_estimator_sharded(work_shard, kT=1.0, group=None, n_total=None):
    """``-kT logsumexp(-w / kT - log n)`` over the union of all ranks' shards (same value on every rank).

    ``n_total`` is the global number of samples; a caller that knows it (as ``bootstrap_statistics_sharded``'s caller
    does) saves the all-reduce of the shard lengths and the host synchronisation that reading its result back costs:
    the call is then asynchronous (one streaming kernel, one 16-byte all-gather, a handful of scalar kernels)."""
    if _world(group) == 1:
        n = work_shard.numel() if n_total is None else int(n_total)
        if work_shard.is_cuda and work_shard.dtype in (torch.float32, torch.float64):
            return _ops.fep_estimate(work_shard, kT, _log_n(n))[0]      # estimate evaluated by the final reduction kernel
    m, s = combine_lse_partials(_ops.lse(work_shard, -1.0 / kT), group)
    if n_total is None:
        n_total = work_shard.numel() if _world(group) == 1 else _total(work_shard.numel(), work_shard.device, group)
    return (-kT * (m + torch.log(s) - _log_n(int(n_total)))).to(work_shard.dtype)


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
    if rng == 'philox':
        seed = int(torch.randint(0, 2**62, (1,), generator=generator).item())
        cells, owner = shard_cells(shard_offset, work_shard.numel(), n_total, dev, group)
        counts = stratified_counts(n_resamples, n_total, cells, n_total, seed)
        rank = dist.get_rank(group) if _world(group) > 1 else 0
        mine = [i for i, o_ in enumerate(owner) if o_ == rank]
        sums = philox_cell_sums(e, shard_offset, cells, mine, counts, seed, n_resamples, n_total)
        sums, refs = repair_underflow(sums, o[0], work_shard, scale, lambda rows, table: philox_cell_sums(
            table, shard_offset, cells, mine, counts, seed, n_resamples, n_total, rows=rows))
    else:
        sums = torch.empty(n_resamples, dtype=torch.float64, device=dev)
        refs = torch.empty(n_resamples, dtype=torch.float64, device=dev)
        gen = torch.default_generator if generator is None else generator
        state = _generator_to_state(gen).to(dev)
        batch = max(1, min(n_resamples if batch is None else batch, (1 << 29) // max(n_total, 1) or 1,
                           _ops.MAX_RESAMPLES_PER_CALL))
        idx = torch.empty(batch * n_total, dtype=torch.int32, device=dev)
        for k in range(0, n_resamples, batch):
            nb = min(batch, n_resamples - k)
            _ops.mt19937_indices(state, nb * n_total, n_total, out=idx)
            rows_idx = idx[:nb * n_total].view(nb, n_total)
            part = _ops.bootstrap_sums(e, n_total, nb, n_total, rows_idx, shard_lo=shard_offset)
            # a resample without a draw in this shard keeps a sum of zero: repair_underflow stops at the shard's minimum
            sums[k:k + nb], refs[k:k + nb] = repair_underflow(part, o[0], work_shard, scale, lambda rows, table: _ops.bootstrap_sums(
                table, n_total, len(rows), n_total, rows_idx[rows].contiguous(), shard_lo=shard_offset))
        _state_to_generator(gen, state)
    sums, ref = combine_bootstrap_sums(sums, refs, group)
    return (-kT * (ref + torch.log(sums) - _log_n(n_total))).to(work_shard.dtype)
