"""The device math headers (tfep_b200/csrc/tx_math.cuh), compiled for the host, against the oracle.

Covers forward, inverse and the hand-derived vector-Jacobian products (vs autograd of the oracle) of every
transformer in fp32 and fp64, plus bin-index exactness of the spline search given identical knots.
"""

import pytest
import torch

import hostcheck as hc
from helpers import rel_err
from oracle import cases
from oracle import flow_oracle as fo

TOL = {torch.float32: 3e-5, torch.float64: 1e-11}


def _run(spec, x, par, inverse=False, **kw):
    if isinstance(spec, fo.Affine):
        return hc.affine(x, par, inverse, **kw)
    if isinstance(spec, fo.Shift):
        return hc.shift(x, par, spec, inverse, **kw)
    if isinstance(spec, fo.SOS):
        return hc.sos(x, par, spec.n_polynomials, **{k: v for k, v in kw.items() if k == 'gy'})
    if isinstance(spec, fo.SymMoebius):
        return hc.moebius(x, par, spec.dimension, spec.max_radius, 2, inverse, **kw)
    if isinstance(spec, fo.Moebius):
        return hc.moebius(x, par, spec.dimension, spec.max_radius, spec.unit_sphere, inverse, **kw)
    return hc.spline(x, par, spec, inverse, **kw)


@pytest.fixture(params=[torch.float32, torch.float64], ids=['f32', 'f64'])
def dtype(request):
    old = torch.get_default_dtype()
    torch.set_default_dtype(request.param)
    yield request.param
    torch.set_default_dtype(old)


def _elementary(dtype):
    return {k: v for k, v in cases.transformer_cases(dtype).items() if not isinstance(v[0], fo.Mixed)}


def test_forward_and_inverse(dtype):
    # Checked against the oracle evaluated in DOUBLE on the same inputs: in the far tails of a spline the
    # fp32 reference itself is only ~1e-4 accurate while the kernels use the
    # analytically identical linear map, so the meaningful target there is the double result.
    for name, (spec, n, x, par) in _elementary(dtype).items():
        y_o, ld_o = cases.double_reference(spec, x, par)
        y, ld = _run(spec, x, par)
        assert rel_err(y, y_o) < TOL[dtype] and rel_err(ld, ld_o) < TOL[dtype], name
        if isinstance(spec, fo.SOS):
            continue
        # inverse: x = T^-1(y) amplifies rounding by 1 / slope (slopes go down to 1e-4), so parity is judged
        # against the same-precision oracle, except in the tails (see above) where the double result rules
        y_in = y_o.to(dtype)
        x_o, ldi_o = cases.double_reference(spec, y_in, par, inverse=True)
        x_p, ldi_p = spec.inverse(y_in, par)
        xi, ldi = _run(spec, y_in, par, inverse=True)
        assert min(rel_err(xi, x_o), rel_err(xi, x_p)) < 3 * TOL[dtype], name
        assert min(rel_err(ldi, ldi_o), rel_err(ldi, ldi_p)) < 3 * TOL[dtype], name


def test_vjp_against_oracle_autograd(dtype):
    for name, (spec, n, x, par) in _elementary(dtype).items():
        gy = cases.normal(tuple(x.shape), 91, dtype)
        gl = cases.normal((x.shape[0],), 92, dtype)
        if isinstance(spec, fo.SOS):
            gx_o, gp_o = spec.vjp(x, par, gy)
        else:
            xg, pg = x.clone().requires_grad_(True), par.clone().requires_grad_(True)
            yy, ll = spec.forward(xg, pg)
            ((yy * gy).sum() + (ll * gl).sum()).backward()
            gx_o, gp_o = xg.grad, pg.grad
        gx, gp = _run(spec, x, par, gy=gy, gl=gl)
        scale = float(1 + gp_o.abs().max())
        assert rel_err(gx, gx_o) < 10 * TOL[dtype], name
        assert float((gp - gp_o).abs().max()) / scale < 10 * TOL[dtype], name


def test_spline_bins_exact_in_double():
    """In fp64 the knots are bit-identical to the oracle's, so every bin index must agree."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        for name, (spec, n, x, par) in cases.transformer_cases(torch.float64).items():
            if not isinstance(spec, fo.Spline):
                continue
            bins_o = fo.spline_bins(spec, x, par)
            _, _, bins = hc.spline(x, par, spec, return_bins=True)
            assert torch.equal(bins.long(), bins_o), name
    finally:
        torch.set_default_dtype(old)


def test_spline_bins_fp32_equal_away_from_knots():
    for name, (spec, n, x, par) in cases.transformer_cases(torch.float32).items():
        if not isinstance(spec, fo.Spline):
            continue
        bins_o = fo.spline_bins(spec, x, par)
        _, _, bins = hc.spline(x, par, spec, return_bins=True)
        assert (bins.long() != bins_o).float().mean() < 0.02, name
