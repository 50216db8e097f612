"""The oracle against the golden vectors generated from the real reference (oracle/make_golden.py)."""

import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden, rel_err
from oracle import analysis_oracle as ao
from oracle import cases
from oracle import flow_oracle as fo

TOL = {'f32': 2e-6, 'f64': 1e-12}
DT = {'f32': torch.float32, 'f64': torch.float64}


@pytest.fixture(params=['f32', 'f64'])
def prec(request):
    old = torch.get_default_dtype()
    torch.set_default_dtype(DT[request.param])
    yield request.param
    torch.set_default_dtype(old)


def test_degree_known_answers():
    tab = json.load(open(os.path.join(GOLDEN, 'degrees.json')))
    for row in tab['generate_degrees']:
        assert fo.gen_degrees(row['n_features'], **row['kwargs']).tolist() == row['expected']
    for row in tab['hidden_degrees']:
        din, dout = torch.tensor(row['degrees_in']), torch.tensor(row['degrees_out'])
        got = fo.hidden_degrees(din, dout, row['hidden_layers'])
        assert [h.tolist() for h in got] == row['expected']
        masks, _ = fo.made_masks(din, dout, row['hidden_layers'])
        assert [int(m.sum()) for m in masks] == row['mask_sums']


def test_transformers_match_golden(prec):
    g = golden(f'transformers_{prec}.npz')
    for name, (spec, n, x, par) in cases.transformer_cases(DT[prec]).items():
        assert np.array_equal(g[f'{name}/x'], x.numpy()) and np.array_equal(g[f'{name}/par'], par.numpy()), name
        y, ld = spec.forward(x, par)
        assert rel_err(y, g[f'{name}/y']) < TOL[prec], name
        assert rel_err(ld, g[f'{name}/ld']) < TOL[prec], name
        if f'{name}/xinv' in g:
            xi, ldi = spec.inverse(torch.from_numpy(g[f'{name}/y']), par)
            assert rel_err(xi, g[f'{name}/xinv']) < 5 * TOL[prec], name
            assert rel_err(ldi, g[f'{name}/ldinv']) < 5 * TOL[prec], name
        if f'{name}/bins' in g:
            bins = fo.spline_bins(spec, x, par)
            assert (bins.numpy() != g[f'{name}/bins']).mean() < 0.01, name
        if isinstance(spec, fo.SOS):
            gx, gp = spec.vjp(x, par, torch.from_numpy(g[f'{name}/gy']))
            assert rel_err(gx, g[f'{name}/gx']) < TOL[prec] and rel_err(gp, g[f'{name}/gpar']) < TOL[prec]


def test_mafs_match_golden(prec):
    g = golden(f'maf_{prec}.npz')
    for name, case in cases.maf_cases(DT[prec]).items():
        oracle, sd = cases.build_oracle(case, DT[prec])
        assert abs(cases.checksum(sd) - float(g[f'{name}/checksum'])) < 1e-9 * float(g[f'{name}/checksum']), name
        y, ld = oracle.forward(case['x'])
        assert rel_err(y, g[f'{name}/y']) < 5 * TOL[prec], name
        assert rel_err(ld, g[f'{name}/ld']) < 5 * TOL[prec], name
        if case['invertible']:
            xi, ldi = oracle.inverse(torch.from_numpy(g[f'{name}/y']))
            assert rel_err(xi, g[f'{name}/xinv']) < 20 * TOL[prec], name
            assert rel_err(ldi, g[f'{name}/ldinv']) < 20 * TOL[prec], name


def test_wrapper_flows_match_golden(prec):
    """oracle/wrappers_oracle.py (Partial / CenteredCentroid / Oriented flows) against the reference's outputs."""
    g = golden(f'wrappers_{prec}.npz')
    for name, case in cases.wrapper_cases(DT[prec]).items():
        oracle, sd = cases.build_wrapper_oracle(case, DT[prec])
        assert abs(cases.checksum(sd) - float(g[f'{name}/checksum'])) < 1e-9 * float(g[f'{name}/checksum']), name
        y, ld = oracle.forward(case['x'].clone())
        assert rel_err(y, g[f'{name}/y']) < 5 * TOL[prec] and rel_err(ld, g[f'{name}/ld']) < 5 * TOL[prec], name
        if case['invertible']:
            xi, ldi = oracle.inverse(torch.from_numpy(g[f'{name}/y']).clone())
            assert rel_err(xi, g[f'{name}/xinv']) < 20 * TOL[prec] and rel_err(ldi, g[f'{name}/ldinv']) < 20 * TOL[prec], name


def test_config_slices_match_golden():
    g = golden('cfg_slices.npz')
    for cfg, nl, B, D in (('cfg1', 2, 64, None), ('cfg2', 4, 64, None), ('cfg3', 6, 32, 30), ('cfg5', 2, 32, 24)):
        flows = cases.cfg_flow(cfg, torch.float32, n_layers=nl, D=D)
        x = cases.cfg_input(cfg, B, torch.float32, D=D)
        assert np.array_equal(x.numpy(), g[f'{cfg}/x'])
        y, ld = fo.sequential([m for m, _ in flows], x)
        assert rel_err(y, g[f'{cfg}/f32/y']) < 2e-5, cfg
        assert rel_err(ld, g[f'{cfg}/f32/ld']) < 2e-5, cfg
        # the fp32 reference itself is within ~1e-5 of the fp64 reference on the same bits
        assert rel_err(g[f'{cfg}/f32/y'], g[f'{cfg}/f64/y']) < 1e-4, cfg


def test_analysis_matches_golden():
    g = golden('analysis.npz')
    w = cases.normal((20000,), 3)
    assert rel_err(ao.fep_estimator(w), g['w_seed3_n20000/fep']) < 1e-6
    assert rel_err(ao.fep_estimator(w, kT=2.5), g['w_seed3_n20000/fep_kT2.5']) < 1e-6
    assert rel_err(ao.fep_estimator(w.double()), g['w_seed3_n20000/fep_f64']) < 1e-12
    wb = torch.stack([w, cases.normal((20000,), 4) * 0.3], dim=1)
    assert rel_err(ao.fep_estimator(wb), g['biased/fep']) < 1e-6
    stats = ao.bootstrap_statistics(w, ao.fep_estimator, 40, batch=11, generator=torch.Generator().manual_seed(1))
    assert rel_err(stats, g['bootstrap/stats_seed1_r40']) < 1e-6
    r = ao.bootstrap(w, ao.fep_estimator, n_resamples=40, batch=9, generator=torch.Generator().manual_seed(1))
    got = [r['confidence_interval']['low'], r['confidence_interval']['high'], r['standard_deviation'], r['mean'], r['median']]
    assert rel_err(torch.stack(got), g['bootstrap/percentile']) < 1e-6


def test_mt19937_restatement_matches_torch_and_golden():
    g = golden('analysis.npz')
    for seed in (0, 1, 12345):
        idx = ao.resample_indices(seed, 3, 700, 20000)
        assert np.array_equal(idx, g[f'randint/seed{seed}_high20000'])
        cont = ao.resample_indices(seed, 2, 700, 100000000, skip=3 * 700)
        assert np.array_equal(cont, g[f'randint/seed{seed}_high1e8_cont'])
        t = torch.randint(0, 20000, (3, 700), generator=torch.Generator().manual_seed(seed)).numpy()
        assert np.array_equal(idx, t)


def test_fep_estimator_known_answer():
    """Work ~ N(0, 1) => Delta f = -1/2 (reference tests/analysis/test_bootstrap.py:178-190)."""
    w = cases.normal((200000,), 17)
    assert abs(float(ao.fep_estimator(w)) + 0.5) < 0.02


def test_kl_loss_matches_golden():
    """oracle kl_loss (loss.py:76-140) on every argument / NaN combination the reference produced."""
    g = golden('analysis.npz')
    u, ld, lw = cases.normal((64,), 5), cases.normal((64,), 6), cases.normal((64,), 7)
    un = u.clone()
    un[[3, 17]] = float('nan')
    assert rel_err(fo.kl_loss(u, ld), g['loss/mean']) == 0
    assert rel_err(fo.kl_loss(u, ld, u * 0.5, lw), g['loss/weighted']) == 0
    for nan in (0, 1):
        for tag, ub in (('clean', u), ('nan', un)):
            for weighted in (0, 1):
                got = fo.kl_loss(ub, ld, u * 0.5, lw if weighted else None, ignore_nan=bool(nan))
                want = torch.as_tensor(g[f'loss/{tag}_w{weighted}_ignore{nan}'])
                assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(want)) and bool(got.isnan() == want.isnan())
