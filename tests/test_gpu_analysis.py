"""Estimator / bootstrap kernels against the oracle, the golden vectors and the reference's index stream."""

import functools

import numpy as np
import pytest
import torch

from helpers import golden, rel_err
from oracle import analysis_oracle as ao
from oracle import cases

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_fep_estimator_against_golden_and_oracle():
    from tfep_b200.analysis import fep_estimator
    g = golden('analysis.npz')
    w = cases.normal((20000,), 3)
    assert rel_err(fep_estimator(w.to(DEV)), g['w_seed3_n20000/fep']) < 1e-6
    assert rel_err(fep_estimator(w.to(DEV), kT=2.5), g['w_seed3_n20000/fep_kT2.5']) < 1e-6
    assert rel_err(fep_estimator(w.double().to(DEV)), g['w_seed3_n20000/fep_f64']) < 1e-12
    wb = torch.stack([w, cases.normal((20000,), 4) * 0.3], dim=1)
    assert rel_err(fep_estimator(wb.to(DEV)), g['biased/fep']) < 1e-6
    for n in (1, 2, 31, 1000, 1 << 20):
        v = cases.normal((n,), 100 + n) * 3
        assert rel_err(fep_estimator(v.to(DEV)), ao.fep_estimator(v.double()).float()) < 2e-6, n
    big = cases.normal((5, 3000), 7)
    assert rel_err(fep_estimator(big.to(DEV), vectorized=True), ao.fep_estimator(big, vectorized=True)) < 1e-6
    wide = torch.tensor([-400.0, 0.0, 300.0, -1e4])        # needs the max-shift: exp would overflow
    assert rel_err(fep_estimator(wide.to(DEV)), ao.fep_estimator(wide.double()).float()) < 1e-6


def test_known_answer_gaussian_work():
    """Work ~ N(0,1) => Delta f = -1/2 (reference tests/analysis/test_bootstrap.py:178-190)."""
    from tfep_b200.analysis import fep_estimator
    w = cases.normal((4000000,), 17).to(DEV)
    assert abs(float(fep_estimator(w)) + 0.5) < 5e-3


def test_resample_indices_bit_exact():
    from tfep_b200 import _ops
    g = golden('analysis.npz')
    for seed in (0, 1, 12345):
        st = _ops.mt19937_seed(seed).to(DEV)
        a = _ops.mt19937_indices(st, 3 * 700, 20000).cpu().numpy().reshape(3, 700)
        assert np.array_equal(a, g[f'randint/seed{seed}_high20000'])
        b = _ops.mt19937_indices(st, 2 * 700, 100000000).cpu().numpy().reshape(2, 700)      # continues the stream
        assert np.array_equal(b, g[f'randint/seed{seed}_high1e8_cont'])
    st = _ops.mt19937_seed(99).to(DEV)
    n = 1000003
    got = torch.cat([_ops.mt19937_indices(st, c, n) for c in (1, 623, 624, 625, 5000, 1248, 77777)]).cpu().numpy()
    assert np.array_equal(got, ao.resample_indices(99, 1, len(got), n)[0])


def test_generator_state_round_trip():
    """The caller's torch.Generator is advanced exactly as the reference's torch.randint calls would."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    w = cases.normal((5000,), 3).to(DEV)
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    bootstrap(w, fep_estimator, n_resamples=7, batch=3, generator=g1)
    ao.bootstrap(w.cpu(), ao.fep_estimator, n_resamples=7, batch=3, generator=g2)
    assert torch.equal(torch.randint(0, 1 << 30, (50,), generator=g1), torch.randint(0, 1 << 30, (50,), generator=g2))


def test_bootstrap_against_golden():
    from tfep_b200.analysis import bootstrap, fep_estimator
    from tfep_b200.analysis.bootstrap import _bootstrap_statistics
    g = golden('analysis.npz')
    w = cases.normal((20000,), 3).to(DEV)
    stats = _bootstrap_statistics(w, fep_estimator, 40, 20000, False, 11, torch.Generator().manual_seed(1), 'mt19937')
    assert rel_err(stats, g['bootstrap/stats_seed1_r40']) < 2e-6
    r = bootstrap(w, fep_estimator, n_resamples=40, batch=9, generator=torch.Generator().manual_seed(1))
    got = torch.stack([r['confidence_interval']['low'], r['confidence_interval']['high'], r['standard_deviation'],
                       r['mean'], r['median']])
    assert rel_err(got, g['bootstrap/percentile']) < 2e-6
    r = bootstrap(w, lambda d, vectorized=False: d.mean(dim=-1), n_resamples=40, batch=7, method='basic',
                  generator=torch.Generator().manual_seed(1))
    got = torch.stack([r['confidence_interval']['low'].reshape(()), r['confidence_interval']['high'].reshape(()),
                       r['standard_deviation'], r['mean'], r['median']])
    assert rel_err(got, g['bootstrap/basic']) < 2e-6
    r = bootstrap(w, fep_estimator, n_resamples=30, bootstrap_sample_size=[100, 5000], take_first_only=True,
                  generator=torch.Generator().manual_seed(2))
    got = torch.stack([torch.stack([x['confidence_interval']['low'], x['confidence_interval']['high'],
                                    x['standard_deviation'], x['mean'], x['median']]) for x in r])
    assert rel_err(got, g['bootstrap/sizes_take_first']) < 5e-6
    kt = bootstrap(w, functools.partial(fep_estimator, kT=2.0), n_resamples=10, generator=torch.Generator().manual_seed(3))
    kt_o = ao.bootstrap(w.cpu(), functools.partial(ao.fep_estimator, kT=2.0), n_resamples=10,
                        generator=torch.Generator().manual_seed(3))
    assert rel_err(kt['mean'], kt_o['mean']) < 2e-6


def test_bootstrap_biased_data_and_errors():
    from tfep_b200.analysis import bootstrap, fep_estimator
    w = torch.stack([cases.normal((3000,), 3), cases.normal((3000,), 4) * 0.3], dim=1)
    a = bootstrap(w.to(DEV), fep_estimator, n_resamples=12, batch=5, generator=torch.Generator().manual_seed(2))
    b = ao.bootstrap(w, ao.fep_estimator, n_resamples=12, batch=5, generator=torch.Generator().manual_seed(2))
    assert rel_err(a['mean'], b['mean']) < 2e-6 and rel_err(a['standard_deviation'], b['standard_deviation']) < 1e-4
    with pytest.raises(ValueError, match='Bayesian bootstrapping does not support'):
        bootstrap(w.to(DEV), fep_estimator, bayesian=True, generator=torch.Generator())


def test_philox_bootstrap_is_statistically_equivalent():
    from tfep_b200.analysis import bootstrap, fep_estimator
    w = cases.normal((200000,), 5).to(DEV)
    R = 2000
    a = bootstrap(w, fep_estimator, n_resamples=R, batch=100, generator=torch.Generator().manual_seed(1))
    b = bootstrap(w, fep_estimator, n_resamples=R, generator=torch.Generator().manual_seed(1), rng='philox')
    # two independent bootstrap means differ by ~ sqrt(2) sigma / sqrt(R); allow 5 of those
    assert abs(float(a['mean']) - float(b['mean'])) < 5 * 2 ** 0.5 * float(a['standard_deviation']) / R ** 0.5
    assert 0.85 < float(b['standard_deviation']) / float(a['standard_deviation']) < 1.18


def test_philox_bootstrap_over_l2_tiles(monkeypatch):
    """Tables larger than the L2 tile are resampled cell by cell: Multinomial counts of draws per cell (host, exact)
    plus uniform draws inside each cell.  Same distribution as uniform draws over the whole table: checked against
    the reference-stream bootstrap; the counts conserve the number of draws."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    import sys
    mod = sys.modules['tfep_b200.analysis.bootstrap']
    w = cases.normal((200000,), 5).to(DEV)
    R = 2000
    a = bootstrap(w, fep_estimator, n_resamples=R, batch=100, generator=torch.Generator().manual_seed(1))
    monkeypatch.setattr(mod, 'L2_TILE_ENTRIES', 30000)
    cells = mod.table_cells(0, 200000)
    assert len(cells) == 7 and cells[0][0] == 0 and cells[-1][1] == 200000
    counts = mod.stratified_counts(50, 200000, cells, 200000, seed=7)
    assert counts.shape == (50, 7) and (counts.sum(axis=1) == 200000).all()
    b = bootstrap(w, fep_estimator, n_resamples=R, generator=torch.Generator().manual_seed(1), rng='philox')
    c = bootstrap(w, fep_estimator, n_resamples=R, generator=torch.Generator().manual_seed(1), rng='philox')
    assert float(b['mean']) == float(c['mean'])                      # deterministic given the generator
    assert abs(float(a['mean']) - float(b['mean'])) < 5 * 2 ** 0.5 * float(a['standard_deviation']) / R ** 0.5
    assert 0.85 < float(b['standard_deviation']) / float(a['standard_deviation']) < 1.18
    # take_first_only sweeps: the table is the first S entries
    d = bootstrap(w, fep_estimator, n_resamples=200, bootstrap_sample_size=[50000, 120000], take_first_only=True,
                  generator=torch.Generator().manual_seed(3), rng='philox')
    e = bootstrap(w, fep_estimator, n_resamples=200, batch=50, bootstrap_sample_size=[50000, 120000], take_first_only=True,
                  generator=torch.Generator().manual_seed(3))
    for x, y in zip(d, e):
        assert abs(float(x['mean']) - float(y['mean'])) < 6 * 2 ** 0.5 * float(y['standard_deviation']) / 200 ** 0.5


def test_bayesian_bootstrap_statistical_parity():
    """Bayesian bootstrap (Dirichlet(1..1) weights; reference bootstrap.py:236-262).  The reference draws the
    weights from the global generator, so parity is statistical: the fused streaming kernel, the generic
    weight-matrix path of this package and the oracle (= reference, pinned in oracle/check_against_reference.py)
    must give the same bootstrap distribution."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    n, R = 4000, 3000
    w = cases.normal((n,), 6)
    torch.manual_seed(11)
    ref = ao.bootstrap(w, ao.fep_estimator, n_resamples=R, bayesian=True)
    torch.manual_seed(12)
    fused = bootstrap(w.to(DEV), fep_estimator, n_resamples=R, bayesian=True)
    generic = bootstrap(w.to(DEV), lambda d, weights=None, vectorized=False: fep_estimator(d, weights=weights, vectorized=vectorized),
                        n_resamples=400, batch=100, bayesian=True)
    sd = float(ref['standard_deviation'])
    for got, r in ((fused, R), (generic, 400)):
        assert abs(float(got['mean']) - float(ref['mean'])) < 5 * 2 ** 0.5 * sd / min(r, R) ** 0.5
        assert 0.8 < float(got['standard_deviation']) / sd < 1.25
    assert abs(float(fused['confidence_interval']['low']) - float(ref['confidence_interval']['low'])) < 0.5 * sd
    assert abs(float(fused['confidence_interval']['high']) - float(ref['confidence_interval']['high'])) < 0.5 * sd
    # sample-size sweep with take_first_only and the argument checks of the reference
    sizes = bootstrap(w.to(DEV), fep_estimator, n_resamples=200, bayesian=True, bootstrap_sample_size=[500, 2000],
                      take_first_only=True)
    assert len(sizes) == 2 and float(sizes[0]['standard_deviation']) > float(sizes[1]['standard_deviation'])
    with pytest.raises(ValueError, match='only when take_first_only'):
        bootstrap(w.to(DEV), fep_estimator, bayesian=True, bootstrap_sample_size=[100])


def test_sharded_partials_add_up_on_one_gpu():
    """Emulate two batch shards on one GPU: per-shard kernels + the host combine == the unsharded result."""
    from tfep_b200 import _ops
    from tfep_b200.analysis.estimator import _log_n, combine_partials
    n, R = 30001, 9
    w = (cases.normal((n,), 5) * 1.5).to(DEV)
    cut = 11111
    parts = torch.stack([_ops.lse(w[:cut], -1.0), _ops.lse(w[cut:], -1.0)])
    m, s = combine_partials(parts)
    df = -(m + torch.log(s) - _log_n(n))
    assert rel_err(df, ao.fep_estimator(w.cpu().double())) < 1e-6
    idx = torch.from_numpy(ao.resample_indices(3, R, n, n)).to(torch.int32).to(DEV)
    total = torch.zeros(R, dtype=torch.float64, device=DEV)
    gmax = parts[:, 0].max()
    for lo, hi in ((0, cut), (cut, n)):
        o = _ops.lse(w[lo:hi], -1.0)
        e = _ops.exp_table(w[lo:hi], -1.0, o[:1])
        total += _ops.bootstrap_sums(e, n, R, n, idx, shard_lo=lo) * torch.exp(o[0] - gmax)
    stats = -(gmax + torch.log(total) - _log_n(n))
    ref = ao.bootstrap_statistics(w.cpu(), ao.fep_estimator, R, generator=torch.Generator().manual_seed(3))
    assert rel_err(stats, ref) < 2e-6


def test_sharded_estimator_single_process_entry_points():
    """fep_estimator_sharded without a process group is the plain estimator, with or without the known total."""
    from tfep_b200.analysis import distributed as D
    from tfep_b200.analysis import fep_estimator
    w = (cases.normal((40001,), 11) * 1.5).to(DEV)
    ref = ao.fep_estimator(w.cpu().double())
    for kT in (1.0, 2.5):
        ref = ao.fep_estimator(w.cpu().double(), kT=kT)
        assert rel_err(D.fep_estimator_sharded(w, kT=kT), ref) < 1e-6
        assert rel_err(D.fep_estimator_sharded(w, kT=kT, n_total=w.numel()), ref) < 1e-6
        assert torch.equal(D.fep_estimator_sharded(w, kT=kT, n_total=w.numel()), fep_estimator(w, kT=kT))


def test_estimator_unaligned_views_and_sizes():
    """Vector loads with scalar head / tail: any offset and length gives the oracle's value."""
    from tfep_b200.analysis import fep_estimator
    base = cases.normal((5000,), 21) * 2
    d = base.to(DEV)
    for off in (0, 1, 2, 3, 5):
        for n in (1, 3, 4, 17, 1023, 4096, 4990):
            v = d[off:off + n]
            assert rel_err(fep_estimator(v), ao.fep_estimator(base[off:off + n].double()).float()) < 2e-6, (off, n)
    d64 = base.double().to(DEV)
    for off in (0, 1):
        assert rel_err(fep_estimator(d64[off:off + 999]), ao.fep_estimator(base[off:off + 999].double())) < 1e-12


@pytest.mark.parametrize('method', ['percentile', 'basic'])
@pytest.mark.parametrize('batch', [None, 100])
def test_generic_statistic_against_scipy(method, batch):
    """The reference's own test (tests/analysis/test_bootstrap.py:test_against_scipy): a generic vectorised statistic
    through the gather path, compared with scipy.stats.bootstrap at the reference's tolerances."""
    import numpy as np
    import scipy.stats
    from tfep_b200.analysis import bootstrap
    data = torch.randn(100, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    ours = bootstrap(data.to(DEV), lambda d, vectorized=True: torch.std(d, dim=-1), n_resamples=10000, batch=batch,
                     method=method, generator=torch.Generator().manual_seed(1))
    ref = scipy.stats.bootstrap(data.unsqueeze(0).numpy(), np.std, vectorized=False, n_resamples=10000, method=method,
                                random_state=np.random.RandomState(0))
    assert abs(float(ours['confidence_interval']['low']) - ref.confidence_interval.low) < 1e-1
    assert abs(float(ours['confidence_interval']['high']) - ref.confidence_interval.high) < 1e-1
    assert abs(float(ours['standard_deviation']) - ref.standard_error) < 1e-2


@pytest.mark.parametrize('rng', ['mt19937', 'philox'])
def test_bootstrap_wide_dynamic_range(rng):
    """Work values spanning hundreds of kT (progressively trained maps, heavy tails): the reference takes a per-row
    logsumexp (bootstrap.py:227-231) and stays finite; the fused table path must too -- with take_first_only the table
    maximum must come from the first samples only, and a resample that misses the dominant samples is recomputed
    against a lower reference (tfep_b200.analysis.bootstrap.repair_underflow)."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    # 1. the early samples lie 1000 kT above the later ones; resamples only see the first 100
    w = torch.cat([torch.full((100,), 1000.0), torch.full((900,), 5.0)]) + cases.normal((1000,), 3) * 0.1
    kw = dict(n_resamples=50, bootstrap_sample_size=[100], take_first_only=True)
    a = bootstrap(w.to(DEV), fep_estimator, generator=torch.Generator().manual_seed(1), rng=rng, **kw)
    b = ao.bootstrap(w, ao.fep_estimator, generator=torch.Generator().manual_seed(1), **kw)
    assert bool(torch.isfinite(a['mean'])) and abs(float(a['mean']) - 1000.0) < 1.0
    if rng == 'mt19937':                  # same index stream: same statistics
        assert rel_err(a['mean'], b['mean']) < 2e-6 and rel_err(a['confidence_interval']['low'], b['confidence_interval']['low']) < 2e-6
    # 2. one dominant outlier 300 kT below everything else: ~37 % of the resamples miss it
    w = cases.normal((2000,), 4) * 0.5
    w[17] = -300.0
    a = bootstrap(w.to(DEV), fep_estimator, n_resamples=200, batch=64, generator=torch.Generator().manual_seed(2), rng=rng)
    b = ao.bootstrap(w, ao.fep_estimator, n_resamples=200, batch=64, generator=torch.Generator().manual_seed(2))
    for k in ('mean', 'median', 'standard_deviation'):
        assert bool(torch.isfinite(a[k])), k
    if rng == 'mt19937':
        assert rel_err(a['mean'], b['mean']) < 5e-6 and rel_err(a['median'], b['median']) < 5e-6
        assert rel_err(a['standard_deviation'], b['standard_deviation']) < 1e-4
    else:                                 # statistical parity: the missing fraction is binomial(200, 0.368)
        from tfep_b200.analysis.bootstrap import _bootstrap_statistics
        st = _bootstrap_statistics(w.to(DEV), fep_estimator, 400, 2000, False, 400, torch.Generator().manual_seed(5), 'philox')
        miss = float((st > -100).float().mean())
        assert 0.28 < miss < 0.46, miss
        assert float(st[st > -100].max()) < 1.0 and float(st[st <= -100].min()) > -300.0 - 1.0


def test_more_than_65535_resamples_per_call():
    """n_resamples beyond the grid.y limit of one launch with the default batch (reference bootstrap.py:126-182)."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    w = cases.normal((50,), 8).to(DEV)
    r = bootstrap(w, fep_estimator, n_resamples=70000, generator=torch.Generator().manual_seed(1))
    assert bool(torch.isfinite(r['mean'])) and float(r['standard_deviation']) > 0


@pytest.mark.parametrize('count,n_streams', [(1, 2), (624 * 3 + 5, 3), (100003, 7), (1 << 20, 148), (3000017, 592), (50000, 4096)])
def test_parallel_mt19937_stream_is_bit_identical(count, n_streams):
    """Jump-ahead sub-streams on all SMs (tfepb_mt19937_indices_parallel) against the single-CTA walk and the numpy
    restatement of torch.randint's CPU stream: same indices, same generator state afterwards, from any position."""
    from tfep_b200 import _ops
    n = 1000003
    for seed, pre in ((42, 0), (7, 1000)):
        a, b = _ops.mt19937_seed(seed).to(DEV), _ops.mt19937_seed(seed).to(DEV)
        if pre:
            _ops.mt19937_indices(a, pre, n, n_streams=1)
            _ops.mt19937_indices(b, pre, n, n_streams=1)
        seq = _ops.mt19937_indices(a, count, n, n_streams=1)
        par = _ops.mt19937_indices(b, count, n, n_streams=n_streams)
        assert torch.equal(seq, par)
        assert np.array_equal(par.cpu().numpy(), ao.resample_indices(seed, 1, count, n, skip=pre)[0])
        # both generators continue identically (the state itself may be a different window of the same sequence)
        assert torch.equal(_ops.mt19937_indices(a, 2000, n, n_streams=1), _ops.mt19937_indices(b, 2000, n, n_streams=1))


def test_parallel_mt19937_is_the_default_for_large_requests_and_fast():
    from tfep_b200 import _ops
    count = 1 << 28
    st = _ops.mt19937_seed(123).to(DEV)
    idx = torch.empty(count, dtype=torch.int32, device=DEV)
    _ops.mt19937_indices(st, count, 100000000, out=idx)             # warm-up: polynomials cached per sub-stream length
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _ops.mt19937_indices(st, count, 100000000, out=idx)
    b.record()
    torch.cuda.synchronize()
    rate = count / (a.elapsed_time(b) * 1e-3)
    print(f'parallel MT19937: {rate / 1e9:.1f} G draws/s')
    assert rate > 3e10
    # the default path of a request of 2^24 draws is the parallel one: compare it with the forced single-CTA walk
    s1, s2 = _ops.mt19937_seed(5).to(DEV), _ops.mt19937_seed(5).to(DEV)
    assert torch.equal(_ops.mt19937_indices(s1, 1 << 24, 77777777), _ops.mt19937_indices(s2, 1 << 24, 77777777, n_streams=1))
    assert torch.equal(_ops.mt19937_indices(s1, 5000, 3), _ops.mt19937_indices(s2, 5000, 3))


def test_generic_statistic_philox_is_seeded_and_vectorized_estimator_is_batched():
    """Generic-statistic route with rng='philox': the caller's generator seeds the device stream (same seed -> same
    result, and the generator advances); fep_estimator(vectorized=True) on many rows (also biased data) equals the
    row-by-row evaluation of the oracle."""
    from tfep_b200.analysis import bootstrap, fep_estimator
    w = cases.normal((3000,), 3).to(DEV)
    mean_stat = lambda d, vectorized=False: d.mean(dim=-1)
    g1, g2 = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    a = bootstrap(w, mean_stat, n_resamples=50, batch=20, generator=g1, rng='philox')
    b = bootstrap(w, mean_stat, n_resamples=50, batch=20, generator=g2, rng='philox')
    c = bootstrap(w, mean_stat, n_resamples=50, batch=20, generator=g2, rng='philox')      # advanced generator
    assert torch.equal(a['mean'], b['mean']) and not torch.equal(b['mean'], c['mean'])
    rows = cases.normal((40, 500), 11)
    assert rel_err(fep_estimator(rows.to(DEV), vectorized=True), ao.fep_estimator(rows, vectorized=True)) < 2e-6
    biased = torch.stack([cases.normal((40, 500), 12), cases.normal((40, 500), 13) * 0.3], dim=-1)
    assert rel_err(fep_estimator(biased.to(DEV), kT=1.5, vectorized=True), ao.fep_estimator(biased, kT=1.5, vectorized=True)) < 2e-6


def test_fused_estimate_entry_point_equals_the_two_step_evaluation():
    """tfepb_fep_estimate (estimate evaluated by the final reduction kernel) against tfepb_lse + the host-side formula of
    estimator.py:75-86, fp32 and fp64 data, with and without log-weights, and against the CPU reference."""
    from tfep_b200 import _ops
    from tfep_b200.analysis.estimator import _log_n
    for dtype, tol in ((torch.float32, 1e-6), (torch.float64, 1e-12)):
        w = (cases.normal((30011,), 21) * 1.7).to(dtype).to(DEV)
        for kT in (1.0, 0.6):
            log_n = _log_n(w.numel())
            res, pair = _ops.fep_estimate(w, kT, log_n)
            o = _ops.lse(w, -1.0 / kT)
            assert torch.equal(pair, o) and res.dtype == dtype and res.dim() == 0
            two_step = (-kT * (o[0] + torch.log(o[1]) - log_n)).to(dtype)
            assert rel_err(res, two_step) <= (1e-7 if dtype == torch.float32 else 1e-15)
            assert rel_err(res, ao.fep_estimator(w.cpu().double(), kT=kT)) < max(tol, 1e-6 if dtype == torch.float32 else 1e-7)
        lw = torch.log_softmax(cases.normal((30011,), 22).to(dtype), dim=0).to(DEV)
        res, _ = _ops.fep_estimate(w, 1.0, 0.0, lw)
        o = _ops.lse(w, -1.0, lw)
        assert rel_err(res, (-(o[0] + torch.log(o[1]))).to(dtype)) <= (1e-7 if dtype == torch.float32 else 1e-15)
