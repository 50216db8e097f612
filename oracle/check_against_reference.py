"""Pin the oracle: compare the restatement with the REAL reference, bit for bit.

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (the build
container).  ``python -m oracle.check_against_reference`` exits non-zero on any
mismatch.  ``tests/test_oracle_vs_reference.py`` runs the same checks under
pytest and skips when the reference is absent (GPU box).
"""

import sys

import torch

from . import analysis_oracle as ao
from . import cases
from . import flow_oracle as fo
from .ref_import import import_reference, reference_available


def to_reference_transformer(ref, spec):
    if isinstance(spec, fo.Affine):
        return ref.AffineTransformer()
    if isinstance(spec, fo.Shift):
        return ref.VolumePreservingShiftTransformer(spec.periodic_indices, spec.periodic_limits)
    if isinstance(spec, fo.Spline):
        return ref.NeuralSplineTransformer(
            x0=spec.x0, xf=spec.xf, n_bins=spec.n_bins, y0=spec.y0, yf=spec.yf, circular=spec.circular,
            identity_boundary_slopes=spec.identity_boundary_slopes, learn_lower_bound=spec.learn_lower_bound,
            learn_upper_bound=spec.learn_upper_bound, min_bin_size=spec.min_bin_size, min_slope=spec.min_slope)
    if isinstance(spec, fo.SOS):
        return ref.SOSPolynomialTransformer(spec.n_polynomials)
    if isinstance(spec, fo.Moebius):
        return ref.MoebiusTransformer(spec.dimension, max_radius=spec.max_radius, unit_sphere=spec.unit_sphere)
    if isinstance(spec, fo.SymMoebius):
        return ref.SymmetrizedMoebiusTransformer(spec.dimension, max_radius=spec.max_radius, identity_eps=spec.identity_eps)
    if isinstance(spec, fo.Mixed):
        return ref.MixedTransformer([to_reference_transformer(ref, t) for t in spec.transformers],
                                    [i.tolist() for i in spec.indices])
    raise TypeError(spec)


def to_reference_maf(ref, case, state_dict):
    emb = case.get('embedding')
    if emb is not None:
        emb = ref.PeriodicEmbedding(emb.n_features_in, emb.limits, emb.periodic_indices)
    maf = ref.MAF(degrees_in=case['degrees_in'], transformer=to_reference_transformer(ref, case['spec']),
                  hidden_layers=case['hidden_layers'], embedding=emb, weight_norm=case['weight_norm'],
                  initialize_identity=False)
    missing, unexpected = maf.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if 'weight' in k or 'bias' in k], missing
    return maf


def _same(a, b, what, fails):
    ok = a.shape == b.shape and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    if not ok:
        err = (a.double() - b.double()).abs().max().item() if a.shape == b.shape else 'shape'
        fails.append(f'{what}: max abs diff {err}')
    return ok


def check_degrees(ref, fails):
    kw_table = [
        (3, {}), (2, dict(order='descending')), (5, dict(max_value=1)), (5, dict(order='descending', max_value=1)),
        (6, dict(conditioning_indices=[0, 3])), (5, dict(order='descending', conditioning_indices=[4])),
        (5, dict(max_value=2, conditioning_indices=[1])),
        (6, dict(order='descending', max_value=2, conditioning_indices=[0, 5])),
        (7, dict(max_value=1, conditioning_indices=[0, 4], repeats=2)),
        (6, dict(order='descending', conditioning_indices=[1, 5], repeats=3)),
        (7, dict(conditioning_indices=[1, 2], repeats=[1, 2, 3])),
        (6, dict(order='descending', max_value=1, conditioning_indices=[2], repeats=[1, 2])),
        (66, {}), (66, dict(order='descending')), (300, dict(repeats=3)), (300, dict(repeats=3, order='descending')),
    ]
    for n, kw in kw_table:
        _same(fo.gen_degrees(n, **kw), ref.generate_degrees(n, **kw), f'gen_degrees({n},{kw})', fails)
    hid_table = [
        ([0, 1, 2], [0, 1, 2], 1), ([0, -1, 1, 2], [0, 1, 2, 3], 2), ([3, 2, 1, -1, 0], [0, 0, 1, 1, 2, 2, 3, 3], 1),
        ([2, -1, 0, 1], [1, 2, 0, 3] * 3, 1), ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [6]),
        ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [6, 4]), ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [[1, 0, -1, 2]]),
        (list(range(66)), list(range(66)) * 25, 2), (list(range(65, -1, -1)), list(range(65, -1, -1)) * 2, 3),
    ]
    for din, dout, hl in hid_table:
        a = fo.hidden_degrees(torch.tensor(din), torch.tensor(dout), hl)
        b = ref.MADE._get_degrees_hidden(torch.tensor(din), torch.tensor(dout), hl)
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f'hidden_degrees({din[:4]}..,{hl})[{i}]', fails)
        m_o, _ = fo.made_masks(torch.tensor(din), torch.tensor(dout), hl)
        made = ref.MADE(torch.tensor(din), torch.tensor(dout), hl)
        for i, mo in enumerate(m_o):
            _same(mo, made.layers[2 * i].mask, f'mask[{i}] of ({din[:4]}..,{hl})', fails)


def check_transformers(ref, dtype, fails):
    for name, (spec, n, x, par) in cases.transformer_cases(dtype).items():
        t = to_reference_transformer(ref, spec)
        y_r, ld_r = t(x, par)
        y_o, ld_o = spec.forward(x, par)
        _same(y_o, y_r, f'{name}/{dtype}/forward y', fails)
        _same(ld_o, ld_r, f'{name}/{dtype}/forward logdet', fails)
        if isinstance(spec, fo.SOS):
            continue
        x_r, ldi_r = t.inverse(y_r, par)
        x_o, ldi_o = spec.inverse(y_r, par)
        _same(x_o, x_r, f'{name}/{dtype}/inverse x', fails)
        _same(ldi_o, ldi_r, f'{name}/{dtype}/inverse logdet', fails)
        torch.manual_seed(5)          # (the symmetrized Moebius identity is a small RANDOM tensor, moebius.py:349-350)
        ident = spec.identity_params(n)
        torch.manual_seed(5)
        _same(ident, t.get_identity_parameters(n), f'{name}/{dtype}/identity', fails)
        deg = fo.gen_degrees(n) if not isinstance(spec, (fo.Moebius, fo.SymMoebius)) else fo.gen_degrees(n, repeats=spec.dimension)
        _same(spec.degrees_out(deg), t.get_degrees_out(deg), f'{name}/{dtype}/degrees_out', fails)


def check_mafs(ref, dtype, fails):
    for name, case in cases.maf_cases(dtype).items():
        oracle, sd = cases.build_oracle(case, dtype)
        maf = to_reference_maf(ref, case, sd)
        with torch.no_grad():
            y_r, ld_r = maf(case['x'])
            y_o, ld_o = oracle.forward(case['x'])
            _same(y_o, y_r, f'maf {name}/{dtype}/forward y', fails)
            _same(ld_o, ld_r, f'maf {name}/{dtype}/forward logdet', fails)
            if case['invertible']:
                x_r, ldi_r = maf.inverse(y_r)
                x_o, ldi_o = oracle.inverse(y_r)
                _same(x_o, x_r, f'maf {name}/{dtype}/inverse x', fails)
                _same(ldi_o, ldi_r, f'maf {name}/{dtype}/inverse logdet', fails)


def to_reference_wrapper(ref, case, state_dict, dtype):
    """The reference's wrapper flows around the reference MAF of a wrapper case (cases.wrapper_cases)."""
    flow = to_reference_maf(ref, case['inner'], state_dict)
    for kind, kw in reversed(case['layers']):
        kw = {k: (torch.tensor(v, dtype=dtype) if k in ('weights', 'origin') else v) for k, v in kw.items()}
        flow = {'partial': ref.PartialFlow, 'centroid': ref.CenteredCentroidFlow, 'oriented': ref.OrientedFlow}[kind](flow, **kw)
    return flow


def check_wrappers(ref, dtype, fails):
    for name, case in cases.wrapper_cases(dtype).items():
        oracle, sd = cases.build_wrapper_oracle(case, dtype)
        flow = to_reference_wrapper(ref, case, sd, dtype)
        with torch.no_grad():
            y_r, ld_r = flow(case['x'].clone())
            y_o, ld_o = oracle.forward(case['x'].clone())
            _same(y_o, y_r, f'wrapper {name}/{dtype}/forward y', fails)
            _same(ld_o, ld_r, f'wrapper {name}/{dtype}/forward logdet', fails)
            if case['invertible']:
                x_r, ldi_r = flow.inverse(y_r.clone())
                x_o, ldi_o = oracle.inverse(y_r.clone())
                _same(x_o, x_r, f'wrapper {name}/{dtype}/inverse x', fails)
                _same(ldi_o, ldi_r, f'wrapper {name}/{dtype}/inverse logdet', fails)


def check_cfg2_slice(ref, fails):
    """Two layers of the headline configuration at a small batch, through SequentialFlow."""
    flows = cases.cfg_flow('cfg2', n_layers=2)
    x = cases.cfg_input('cfg2', 32)
    mafs = []
    for m, sd in flows:
        case = dict(degrees_in=m.degrees_in, spec=m.transformer, hidden_layers=2, weight_norm=True)
        mafs.append(to_reference_maf(ref, case, sd))
    seq = ref.SequentialFlow(*mafs)
    with torch.no_grad():
        y_r, ld_r = seq(x)
        y_o, ld_o = fo.sequential([m for m, _ in flows], x)
        _same(y_o, y_r, 'cfg2 sequential forward y', fails)
        _same(ld_o, ld_r, 'cfg2 sequential forward logdet', fails)


def check_analysis(ref, fails):
    w = cases.normal((5000,), 3)
    _same(ao.fep_estimator(w), ref.fep_estimator(w), 'fep_estimator', fails)
    _same(ao.fep_estimator(w.double(), kT=2.5), ref.fep_estimator(w.double(), kT=2.5), 'fep_estimator f64 kT', fails)
    wb = torch.stack([w, cases.normal((5000,), 4) * 0.3], dim=1)
    _same(ao.fep_estimator(wb), ref.fep_estimator(wb), 'fep_estimator biased', fails)
    for kw in [dict(n_resamples=50, batch=7), dict(n_resamples=33), dict(n_resamples=20, bootstrap_sample_size=[100, 1000]),
               dict(n_resamples=20, bootstrap_sample_size=[500], take_first_only=True, method='percentile')]:
        a = ao.bootstrap(w, ao.fep_estimator, generator=torch.Generator().manual_seed(1), **kw)
        b = ref.bootstrap(w, ref.fep_estimator, generator=torch.Generator().manual_seed(1), **kw)
        a, b = (a if isinstance(a, list) else [a]), (b if isinstance(b, list) else [b])
        for i, (ra, rb) in enumerate(zip(a, b)):
            for k in ('standard_deviation', 'mean', 'median'):
                _same(ra[k], rb[k], f'bootstrap {kw} [{i}] {k}', fails)
            for k in ('low', 'high'):
                _same(ra['confidence_interval'][k], rb['confidence_interval'][k], f'bootstrap {kw} [{i}] ci {k}', fails)
    # Bayesian bootstrap: Dirichlet weights from the global generator
    for kw in [dict(n_resamples=15, batch=4), dict(n_resamples=9, bootstrap_sample_size=[300], take_first_only=True)]:
        torch.manual_seed(5)
        a = ao.bootstrap(w, ao.fep_estimator, bayesian=True, **kw)
        torch.manual_seed(5)
        b = ref.bootstrap(w, ref.fep_estimator, bayesian=True, **kw)
        for k in ('standard_deviation', 'mean', 'median'):
            _same(a[k], b[k], f'bayesian bootstrap {kw} {k}', fails)
        for k in ('low', 'high'):
            _same(a['confidence_interval'][k], b['confidence_interval'][k], f'bayesian bootstrap {kw} ci {k}', fails)
    a = ao.bootstrap(wb, ao.fep_estimator, n_resamples=12, batch=5, generator=torch.Generator().manual_seed(2))
    b = ref.bootstrap(wb, ref.fep_estimator, n_resamples=12, batch=5, generator=torch.Generator().manual_seed(2))
    _same(a['mean'], b['mean'], 'bootstrap biased mean', fails)
    # loss
    u, ld = cases.normal((64,), 5), cases.normal((64,), 6)
    _same(fo.kl_loss(u, ld), ref.BoltzmannKLDivLoss()(u, ld), 'kl loss', fails)
    lw, ua = cases.normal((64,), 7), cases.normal((64,), 8)
    un = u.clone()
    un[[3, 17]] = float('nan')
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # the reference calls softmax without dim
        for nan in (False, True):
            L = ref.BoltzmannKLDivLoss(ignore_nan=nan)
            _same(fo.kl_loss(un, ld, ignore_nan=nan), L(un, ld), f'kl loss nan={nan}', fails)
            _same(fo.kl_loss(un, ld, ua, lw, ignore_nan=nan), L(un, ld, log_weights=lw, ref_potentials=ua), f'kl loss weighted nan={nan}', fails)
            _same(fo.kl_loss(u, None, None, lw, ignore_nan=nan), L(u, log_weights=lw), f'kl loss only weights nan={nan}', fails)


def run_all():
    ref = import_reference()
    fails = []
    check_degrees(ref, fails)
    old = torch.get_default_dtype()
    for dtype in (torch.float32, torch.float64):
        torch.set_default_dtype(dtype)      # the reference's own tests run under a double default
        try:
            check_transformers(ref, dtype, fails)
            check_mafs(ref, dtype, fails)
            check_wrappers(ref, dtype, fails)
        finally:
            torch.set_default_dtype(old)
    check_cfg2_slice(ref, fails)
    check_analysis(ref, fails)
    return fails


if __name__ == '__main__':
    if not reference_available():
        print('reference not available; nothing checked')
        sys.exit(0)
    f = run_all()
    for line in f:
        print('MISMATCH', line)
    print('oracle == reference: OK' if not f else f'{len(f)} mismatches')
    sys.exit(1 if f else 0)
