"""Bootstrap distribution of a statistic (reference tfep/analysis/bootstrap.py:24-262).

Same signature and result dictionary as the reference.  Resample indices follow the reference's
contract -- ``torch.randint`` on a CPU generator, i.e. ``MT19937(seed).u32 % max_idx`` row-major over
``(n_resamples, sample_size)``, continuing across ``batch`` chunks (bootstrap.py:207-218) -- and are
produced on the device by tfepb_mt19937_indices, bit for bit, advancing the caller's generator exactly
as the reference would.  ``rng='philox'`` is a counter-based alternative (statistical parity only).

When ``statistic`` is :func:`tfep_b200.analysis.fep_estimator` on 1-D work values the resample + statistic
loop is fused (tfepb_exp_table once, then tfepb_bootstrap_sums: gather + add, no ``batch x n`` temporary);
any other statistic goes through gather + the user's vectorised callable.
"""

import functools
import struct

import numpy as np
import torch

from .. import _ops
from .estimator import _log_n, fep_estimator, lse_partial

_STATE_OFFSET, _N = 24, 624


def _generator_to_state(generator):
    """at::mt19937 state of a CPU generator as 625 int32 words (624 state words + position)."""
    raw = generator.get_state().numpy().tobytes()
    _, left, _, _ = struct.unpack_from('<QiiQ', raw, 0)
    words = np.frombuffer(raw, dtype=np.uint64, count=_N, offset=_STATE_OFFSET).astype(np.uint32)
    st = np.empty(_N + 1, dtype=np.uint32)
    st[:_N] = words
    st[_N] = _N - (left - 1)
    return torch.from_numpy(st.view(np.int32).copy())


def _state_to_generator(generator, state):
    """Write the advanced MT19937 state back so the generator continues where the kernel stopped."""
    st = state.cpu().numpy().view(np.uint32)
    raw = bytearray(generator.get_state().numpy().tobytes())
    pos = int(st[_N])
    seed, _, seeded, _ = struct.unpack_from('<QiiQ', raw, 0)
    struct.pack_into('<QiiQ', raw, 0, seed, _N - pos + 1, seeded, pos)
    raw[_STATE_OFFSET:_STATE_OFFSET + 8 * _N] = st[:_N].astype(np.uint64).tobytes()
    generator.set_state(torch.frombuffer(raw, dtype=torch.uint8).clone())


def _fused_kT(statistic):
    """kT if ``statistic`` is (a partial of) this package's fep_estimator, else None."""
    if statistic is fep_estimator:
        return 1.0
    if isinstance(statistic, functools.partial) and statistic.func is fep_estimator and not statistic.args:
        extra = set(statistic.keywords) - {'kT'}
        if not extra:
            return float(statistic.keywords.get('kT', 1.0))
    return None


def bootstrap(
        data, statistic, *,
        confidence_level=0.95,
        n_resamples=9999,
        bootstrap_sample_size=None,
        take_first_only=False,
        batch=None,
        method='percentile',
        bayesian=False,
        generator=None,
        rng='mt19937',
):
    """Compute confidence interval, standard deviation, mean and median of the bootstrap distribution of
    ``statistic``.  See the reference docstring (bootstrap.py:35-125) for the arguments; ``rng`` is the only
    addition (``'mt19937'``: the reference's index stream, ``'philox'``: counter-based, not bit-compatible).

    Returns a dict (or a list of dicts, one per ``bootstrap_sample_size``) with keys ``confidence_interval``
    (``low`` / ``high``), ``standard_deviation``, ``mean`` and ``median``.
    """
    n_samples = len(data)
    if bayesian and generator is not None:
        raise ValueError('Bayesian bootstrapping does not support random number generators.')
    if bootstrap_sample_size is None:
        bootstrap_sample_size = [n_samples]
    elif bayesian and not take_first_only:
        raise ValueError('With Bayesian bootstrapping, specifying a bootstrap_sample_size '
                         'is supported only when take_first_only is True.')
    if rng not in ('mt19937', 'philox'):
        raise ValueError("rng must be 'mt19937' or 'philox'")
    if batch is None:
        batch = n_resamples

    with torch.no_grad():
        results = []
        for sample_size in bootstrap_sample_size:
            if bayesian:
                stats = _bayesian_bootstrap_statistics(data[:sample_size], statistic, n_resamples, batch)
            else:
                stats = _bootstrap_statistics(data, statistic, n_resamples, sample_size, take_first_only, batch,
                                              generator, rng)
            alpha = (1 - confidence_level) / 2
            quantiles = torch.tensor([alpha, 1 - alpha], dtype=stats.dtype, device=stats.device)
            ci_l, ci_u = torch.quantile(stats, q=quantiles)
            if method == 'basic':
                full_statistic = statistic(data.unsqueeze(0))
                ci_l, ci_u = 2 * full_statistic - ci_u, 2 * full_statistic - ci_l
            results.append(dict(
                confidence_interval=dict(low=ci_l, high=ci_u),
                standard_deviation=torch.std(stats),
                mean=torch.mean(stats),
                median=torch.median(stats),
            ))
    if len(bootstrap_sample_size) == 1:
        return results[0]
    return results


#: Entries of the exp table one pass works on: 32 MB of fp32, comfortably resident in the 126 MB L2 of a B200.
#: Measured per 1000 resamples of 1e8 draws (scripts/dev_bootstrap_tiles.py): 1.55 s with the whole 400 MB table as
#: one cell (gathers from HBM), 0.85 s with 128 MB cells, 0.35-0.36 s for every cell size from 8 to 64 MB.
L2_TILE_ENTRIES = 8 << 20


def table_cells(lo, hi, tile=None):
    """Cut the index range [lo, hi) into equal cells of at most ``tile`` entries: list of (begin, end)."""
    tile = L2_TILE_ENTRIES if tile is None else tile
    n = hi - lo
    k = max(1, -(-n // tile))
    b = [lo + (n * i) // k for i in range(k + 1)]
    return list(zip(b[:-1], b[1:]))


def stratified_counts(n_resamples, sample_size, cells, total, seed):
    """How many of the ``sample_size`` uniform draws over [0, total) of every resample fall into each cell: exact
    Multinomial(sample_size; cell sizes / total) samples, (n_resamples, n_cells) int64 on the host.  A resample of
    uniform draws IS such a count vector plus that many uniform draws inside each cell, so the kernels can work
    cell by cell -- an L2-resident tile of the table at a time, or only the cells a rank owns -- without
    generating any draw twice.  Same seed -> same counts on every rank."""
    if len(cells) == 1:
        return None
    p = np.array([(b - a) / total for a, b in cells], dtype=np.float64)
    rng = np.random.Generator(np.random.Philox(key=int(seed) & (2**64 - 1)))
    return rng.multinomial(int(sample_size), p / p.sum(), size=int(n_resamples)).astype(np.int64)


def philox_cell_sums(e, e_lo, cells, mine, counts, seed, n_resamples=None, sample_size=None, rows=None):
    """Per-resample sums over the cells ``mine`` (indices into ``cells``) of the exp table ``e``, which holds the
    entries [e_lo, e_lo + len(e)) of the global table.  Returns (n_resamples,) float64, or, with ``rows`` (a list of
    resample numbers), the sums of those resamples only -- the same Philox counters, so the same draws."""
    if counts is None:                                   # a single cell: plain uniform draws over it
        (a, b), = cells
        stride = (sample_size + 3) // 4
        if rows is not None:
            return torch.cat([_ops.bootstrap_sums(e[a - e_lo:b - e_lo], b - a, 1, sample_size, None, seed, r * stride)
                              for r in rows]) if len(rows) else torch.empty(0, dtype=torch.float64, device=e.device)
        sums = torch.empty(n_resamples, dtype=torch.float64, device=e.device)
        for k in range(0, n_resamples, 65535):
            nb = min(65535, n_resamples - k)
            sums[k:k + nb] = _ops.bootstrap_sums(e[a - e_lo:b - e_lo], b - a, nb, sample_size, None, seed, k * stride)
        return sums
    n_resamples = counts.shape[0]
    strides = (counts.max(axis=0) + 3) // 4               # Philox counters: a disjoint range per (cell, resample)
    offsets = np.concatenate([[0], np.cumsum(strides * n_resamples)])
    if rows is not None:
        out = torch.zeros(len(rows), dtype=torch.float64, device=e.device)
        for c in mine:
            a, b = cells[c]
            for i, r in enumerate(rows):
                if counts[r, c] > 0:
                    size = torch.tensor([int(counts[r, c])], dtype=torch.int64, device=e.device)
                    out[i:i + 1] += _ops.bootstrap_sums(e[a - e_lo:b - e_lo], b - a, 1, int(counts[:, c].max()), None, seed,
                                                        int(offsets[c]) + r * int(strides[c]), sample_sizes=size)
        return out
    sums = torch.zeros(n_resamples, dtype=torch.float64, device=e.device)
    for c in mine:
        a, b = cells[c]
        sizes = torch.from_numpy(np.ascontiguousarray(counts[:, c])).to(e.device)
        for k in range(0, n_resamples, 65535):
            nb = min(65535, n_resamples - k)
            sums[k:k + nb] += _ops.bootstrap_sums(e[a - e_lo:b - e_lo], b - a, nb, int(counts[:, c].max()), None, seed,
                                                  int(offsets[c]) + k * int(strides[c]), sample_sizes=sizes[k:k + nb])
    return sums


#: A resample whose sum of exp(v - reference) falls below this is recomputed against a lower reference (see
#: repair_underflow): its draws all lie more than 69 below the reference, where the fp32 table runs out of range.
UNDERFLOW = 1e-30
LEVEL_STEP = 60.0       # < -log(UNDERFLOW): the entries such a resample drew stay below exp(-9) at the next level


def repair_underflow(sums, ref0, values, scale, rerun):
    """Make the fused path exact for data of ANY dynamic range.  ``sums[r] = sum_j exp(v[idx_rj] - ref0)`` comes from an
    fp32 table that underflows for entries more than ~87 below the reference; the reference's per-row logsumexp
    (bootstrap.py:227-231 with estimator.py:84-86) does not.  A resample that missed every dominant sample -- e.g. all
    of them with ``take_first_only`` on progressively trained work values, or a resample of a heavy-tailed data set
    that skips the outlier -- ends up with a sum of (almost) zero: those rows are recomputed with the same draws
    against tables at references lowered in steps of 60 until they register (entries above a lowered reference
    overflow to inf, but a row that needs the level drew none of them).  ``rerun(rows, table)`` returns the sums of
    the resamples ``rows`` over ``table``.  Returns ``(sums, refs)``: per-row sums and the reference each is relative to.
    """
    refs = torch.full_like(sums, float(ref0))
    bad = sums < UNDERFLOW
    if not bool(bad.any()):
        return sums, refs
    vmin = float((values * scale).min())
    ref = float(ref0)
    while bool(bad.any()) and ref > vmin + LEVEL_STEP:
        ref -= LEVEL_STEP
        rows = bad.nonzero().flatten()
        table = _ops.exp_table(values, scale, torch.tensor([ref], dtype=torch.float64, device=values.device))
        sums[rows] = rerun(rows.tolist(), table)
        refs[rows] = ref
        bad = sums < UNDERFLOW
    return sums, refs


def bootstrap_partial_sums(data, kT, n_resamples, sample_size, max_idx, batch, generator, rng,
                           shard_offset=0, global_max=None):
    """Per-resample ``sum_j exp(v[idx_rj] - ref_r)`` over the draws of every resample (``v = -data / kT``).

    The table the draws index is ``data[:max_idx]`` -- with ``take_first_only`` the resamples only see the first
    ``bootstrap_sample_size`` samples (reference bootstrap.py:207-218), so its reference maximum must not come from
    later ones.  Returns ``(sums (n_resamples,) float64, refs (n_resamples,) float64)``; ``refs`` is the table maximum
    except for rows repaired by :func:`repair_underflow`.  Single GPU: shard = all; the sharded version lives in
    tfep_b200.analysis.distributed.
    """
    if shard_offset != 0 or global_max is not None:
        raise NotImplementedError('sharded bootstrap: see tfep_b200.analysis.distributed')
    scale = -1.0 / kT
    values = data[:max_idx]
    o = _ops.lse(values, scale)
    e = _ops.exp_table(values, scale, o[:1])
    if rng == 'philox':
        seed = int(torch.randint(0, 2**62, (1,), generator=generator).item())
        cells = table_cells(0, max_idx)
        counts = stratified_counts(n_resamples, sample_size, cells, max_idx, seed)
        everything = range(len(cells))
        sums = philox_cell_sums(e, 0, cells, everything, counts, seed, n_resamples, sample_size)
        return repair_underflow(sums, o[0], values, scale, lambda rows, table: philox_cell_sums(
            table, 0, cells, everything, counts, seed, n_resamples, sample_size, rows=rows))
    sums = torch.empty(n_resamples, dtype=torch.float64, device=data.device)
    refs = torch.empty(n_resamples, dtype=torch.float64, device=data.device)
    gen = torch.default_generator if generator is None else generator
    state = _generator_to_state(gen).to(data.device)
    batch = max(1, min(batch, n_resamples, _ops.MAX_RESAMPLES_PER_CALL))
    idx = torch.empty(batch * sample_size, dtype=torch.int32, device=data.device)
    for k in range(0, n_resamples, batch):
        nb = min(batch, n_resamples - k)
        _ops.mt19937_indices(state, nb * sample_size, max_idx, out=idx)
        rows_idx = idx[:nb * sample_size].view(nb, sample_size)
        part = _ops.bootstrap_sums(e, max_idx, nb, sample_size, rows_idx)
        sums[k:k + nb], refs[k:k + nb] = repair_underflow(part, o[0], values, scale, lambda rows, table: _ops.bootstrap_sums(
            table, max_idx, len(rows), sample_size, rows_idx[rows].contiguous()))
    _state_to_generator(gen, state)
    return sums, refs


def _bootstrap_statistics(data, statistic, n_resamples, sample_size, take_first_only, batch, generator, rng):
    """The ``n_resamples`` values of the statistic (reference bootstrap.py:185-233)."""
    n_samples = len(data)
    max_idx = sample_size if take_first_only else n_samples
    kT = _fused_kT(statistic)
    if kT is not None and data.dim() == 1:
        batch = max(1, min(batch, n_resamples, (1 << 29) // max(sample_size, 1) or 1))
        sums, refs = bootstrap_partial_sums(data, kT, n_resamples, sample_size, max_idx, batch, generator, rng)
        return (-kT * (refs + torch.log(sums) - _log_n(sample_size))).to(data.dtype)

    # generic statistic: indices from the same stream, gather, user callable on (batch, sample_size[, dim])
    gen = torch.default_generator if generator is None else generator
    stats = torch.empty(n_resamples, dtype=data.dtype, device=data.device)
    state = _generator_to_state(gen).to(data.device) if rng == 'mt19937' else None
    dev_gen = None
    if rng == 'philox':        # the caller's generator seeds the device stream: reproducible, and its state advances
        dev_gen = torch.Generator(device=data.device).manual_seed(int(torch.randint(0, 2**62, (1,), generator=generator).item()))
    for k in range(0, n_resamples, batch):
        nb = min(batch, n_resamples - k)
        if rng == 'mt19937':
            idx = _ops.mt19937_indices(state, nb * sample_size, max_idx).view(nb, sample_size).long()
        else:
            idx = torch.randint(0, max_idx, (nb, sample_size), device=data.device, generator=dev_gen)
        expanded = data.expand((nb, *data.shape))
        if data.dim() > 1:
            idx = idx.unsqueeze(-1).expand(nb, sample_size, data.shape[1])
        stats[k:k + nb] = statistic(torch.gather(expanded, dim=1, index=idx), vectorized=True)
    if state is not None:
        _state_to_generator(gen, state)
    return stats


def _bayesian_bootstrap_statistics(data, statistic, n_resamples, batch):
    """Statistics under Dirichlet(1, ..., 1) sample weights (reference bootstrap.py:236-262).

    The reference draws the weights from the global generator (``Dirichlet.sample``; no generator argument), so
    only statistical parity is defined.  With this package's ``fep_estimator`` on 1-D work values the weights are
    never materialised: Dirichlet(1..1) = normalised Exp(1) variates g, and
    ``-kT logsumexp(v + log(g / sum g)) = -kT (log sum_i e^{v_i} g_i - log sum_i g_i)`` is one streaming pass per
    resample (tfepb_bayesian_bootstrap_sums).  Any other statistic receives a ``(batch, n)`` weight matrix."""
    n = len(data)
    kT = _fused_kT(statistic)
    if kT is not None and data.dim() == 1:
        scale = -1.0 / kT
        o = _ops.lse(data, scale)
        e = _ops.exp_table(data, scale, o[:1])
        seed = int(torch.randint(0, 2**62, (1,)).item())
        stats = torch.empty(n_resamples, dtype=torch.float64, device=data.device)
        for k in range(0, n_resamples, 65535):
            nb = min(65535, n_resamples - k)
            s, g = _ops.bayesian_bootstrap_sums(e, nb, seed, k * ((n + 3) // 4))
            stats[k:k + nb] = -kT * (o[0] + torch.log(s) - torch.log(g))
        return stats.to(data.dtype)
    stats = torch.empty(n_resamples, dtype=data.dtype, device=data.device)
    dirichlet = torch.distributions.Dirichlet(torch.ones(n, dtype=data.dtype, device=data.device))
    for k in range(0, n_resamples, batch):
        nb = min(batch, n_resamples - k)
        weights = dirichlet.sample((nb,))
        stats[k:k + nb] = statistic(data.expand((nb, *data.shape)), weights=weights, vectorized=True)
    return stats
