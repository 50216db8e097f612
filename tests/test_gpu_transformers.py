"""CUDA transformer kernels (through the C ABI) against the oracle and the reference's golden vectors."""

import numpy as np
import pytest
import torch

from helpers import golden, rel_err, to_module
from oracle import cases
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = {'f32': 1e-5, 'f64': 1e-10}        # north_star: 1e-5 relative for the fp32 path
DT = {'f32': torch.float32, 'f64': torch.float64}


@pytest.fixture(params=['f32', 'f64'])
def prec(request):
    old = torch.get_default_dtype()
    torch.set_default_dtype(DT[request.param])
    yield request.param
    torch.set_default_dtype(old)


def test_forward_inverse_against_golden(prec):
    g = golden(f'transformers_{prec}.npz')
    g64 = golden('transformers_f64.npz')
    for name, (spec, n, x, par) in cases.transformer_cases(DT[prec]).items():
        mod = to_module(spec).to(DEV)
        y, ld = mod(x.to(DEV), par.to(DEV))
        # Non-circular splines with inputs in the far tails: the fp32 reference is itself only ~1e-4 accurate
        # there (it subtracts knots placed at 1000 x the domain width); the kernels
        # evaluate the analytically identical linear map, so those cases are held to the reference's DOUBLE
        # result (same inputs up to fp32 rounding) at the same 1e-5.
        tails = prec == 'f32' and f'{name}/bins' in g and bool(((g[f'{name}/bins'] == 0) |
                                                                (g[f'{name}/bins'] == spec.n_bins + 1)).any())
        ref = g64 if tails else g
        assert rel_err(y, ref[f'{name}/y']) < TOL[prec], name
        assert rel_err(ld, ref[f'{name}/ld']) < TOL[prec], name
        if isinstance(spec, fo.SOS):
            with pytest.raises(NotImplementedError):
                mod.inverse(y, par.to(DEV))
            continue
        xi, ldi = mod.inverse(torch.from_numpy(g[f'{name}/y']).to(DEV), par.to(DEV))
        # The inverse amplifies the last-ulp differences of exp / log between hosts and devices by 1 / slope
        # (slopes go down to min_slope = 1e-4 in these cases): 2e-3 in fp32, still 5e-10 in fp64.
        inv_tol = 2e-3 if prec == 'f32' and isinstance(spec, (fo.Spline, fo.Mixed)) else 5 * TOL[prec]
        assert rel_err(xi, g[f'{name}/xinv']) < inv_tol, name
        assert rel_err(ldi, g[f'{name}/ldinv']) < inv_tol, name
        if tails:
            xi64, ldi64 = cases.double_reference(spec, torch.from_numpy(g[f'{name}/y']), par, inverse=True)
            # (the fp32 golden y fed to the inverse carries the reference's own ~1e-4 tail error)
            assert rel_err(xi, xi64) < 20 * TOL[prec] and rel_err(ldi, ldi64) < 20 * TOL[prec], name


def test_round_trip(prec):
    for name, (spec, n, x, par) in cases.transformer_cases(DT[prec]).items():
        if isinstance(spec, fo.SOS):
            continue
        mod = to_module(spec).to(DEV)
        y, ld = mod(x.to(DEV), par.to(DEV))
        xi, ldi = mod.inverse(y, par.to(DEV))
        xx = x
        if isinstance(spec, fo.Spline) and spec.circular:      # periodic: compare modulo the period
            period = (spec.xf - spec.x0)
            d = (xi.cpu() - x).abs()
            # x = T^-1(y) amplifies the rounding of y by 1 / (dy/dx), and slopes go down to min_slope = 1e-4
            assert float(torch.minimum(d, (period - d).abs()).max()) < (5e-3 if prec == 'f32' else 1e-8), name
        elif isinstance(spec, fo.Shift) and spec.periodic_indices is not None:
            # the reference wraps as `v % P + lower` in both directions, which is not its own inverse: the
            # non-periodic features must come back, the periodic ones must equal the oracle's inverse
            keep = [i for i in range(n) if i not in spec.periodic_indices.tolist()]
            assert rel_err(xi[:, keep], xx[:, keep]) < 50 * TOL[prec], name
            assert rel_err(xi, spec.inverse(y.cpu(), par)[0]) < 50 * TOL[prec], name
        elif not isinstance(spec, fo.Mixed):
            assert rel_err(xi, xx) < 50 * TOL[prec], name
        # log-dets of +-16 from slopes near min_slope: the fp32 round trip cancels to ~1e-3 absolute
        assert rel_err(ld + ldi, torch.zeros_like(ld)) < (5e-3 if prec == 'f32' else 1e-8), name


def test_spline_bin_indices(prec):
    """fp64: every index equals the reference's.  fp32: equal except within a few ulp of a knot."""
    g = golden(f'transformers_{prec}.npz')
    for name, (spec, n, x, par) in cases.transformer_cases(DT[prec]).items():
        if not isinstance(spec, fo.Spline):
            continue
        bins = to_module(spec).to(DEV).bin_indices(x.to(DEV), par.to(DEV)).cpu().numpy()
        ref = g[f'{name}/bins']
        if prec == 'f64':
            assert np.array_equal(bins, ref), name
        else:
            assert (bins != ref).sum() <= 1, name           # at most one element on a knot (72-144 elements per case)


def test_spline_bin_indices_one_million_elements():
    """north_star: bit-exact spline bin indices.  1.08 M (sample, feature) pairs with the parameters cfg2's first
    conditioner produces (CPU oracle, fp32) go through the CUDA kernel and the reference's `sum(x > knots) - 1`
    (spline.py:622-625): the indices must be identical except where the wrapped input lies within 4 ulp of a knot (the
    kernel's wrap `(x - x0 + shift) mod L` and the reference's differ by at most an ulp there) -- and those are counted."""
    flows = cases.cfg_flow('cfg2', n_layers=1)
    oracle, _ = flows[0]
    spec = oracle.transformer
    B = 16384
    x = cases.cfg_input('cfg2', B)
    with torch.no_grad():
        par = oracle.parameters_of(x)
        ref = fo.spline_bins(spec, x, par).numpy()
        x0, y0, w, h, d, shifts = spec.unpack(par)
        xw = ((x - x0 + shifts) % (spec.xf - x0) + x0).numpy()
        knots = np.concatenate([np.broadcast_to(x0.numpy(), (B, 1, x.shape[1])),
                                (x0.unsqueeze(-2) + torch.cumsum(w, dim=1)).numpy()], axis=1)      # (B, K + 1, F)
    bins = to_module(spec).to(DEV).bin_indices(x.to(DEV), par.to(DEV)).cpu().numpy()
    assert bins.shape == ref.shape and bins.size >= 1000000
    bad = np.argwhere(bins != ref)
    for b, f in bad:
        dist = np.abs(knots[b, :, f] - xw[b, f]).min()
        assert abs(int(bins[b, f]) - int(ref[b, f])) == 1, (b, f, bins[b, f], ref[b, f])
        assert dist <= 4 * np.spacing(np.float32(max(abs(xw[b, f]), np.pi))), (b, f, dist)
    assert len(bad) <= 20, len(bad)                        # ~1e-5 of the pairs at most sit that close to a knot
    print(f'bin indices: {len(bad)} of {bins.size} differ, all within 4 ulp of a knot')


def test_vjp_against_oracle_autograd(prec):
    dtype = DT[prec]
    for name, (spec, n, x, par) in cases.transformer_cases(dtype).items():
        gy = cases.normal(tuple(x.shape), 91, dtype)
        gl = cases.normal((x.shape[0],), 92, dtype)
        xg, pg = x.clone().requires_grad_(True), par.clone().requires_grad_(True)
        yy, ll = spec.forward(xg, pg)
        if isinstance(spec, fo.SOS):
            gx_o, gp_o = spec.vjp(x, par, gy)
        else:
            ((yy * gy).sum() + (ll * gl).sum()).backward()
            gx_o, gp_o = xg.grad, pg.grad
        xd, pd = x.to(DEV).requires_grad_(True), par.to(DEV).requires_grad_(True)
        y, ld = to_module(spec).to(DEV)(xd, pd)
        loss = (y * gy.to(DEV)).sum()
        if ld.requires_grad:
            loss = loss + (ld * gl.to(DEV)).sum()
        else:
            assert isinstance(spec, (fo.SOS, fo.Shift))   # the reference's SOS log-det carries no gradient; the
                                                          # volume-preserving shift returns constant zeros
        loss.backward()
        scale = float(1 + gp_o.abs().max())
        assert rel_err(xd.grad, gx_o) < 20 * TOL[prec], name
        assert float((pd.grad.cpu() - gp_o).abs().max()) / scale < 20 * TOL[prec], name


def test_sos_backward_matches_reference_golden(prec):
    g = golden(f'transformers_{prec}.npz')
    for name in ('sos2', 'sos3'):
        spec, n, x, par = cases.transformer_cases(DT[prec])[name]
        xd, pd = x.to(DEV).requires_grad_(True), par.to(DEV).requires_grad_(True)
        y, _ = to_module(spec).to(DEV)(xd, pd)
        (y * torch.from_numpy(g[f'{name}/gy']).to(DEV)).sum().backward()
        assert rel_err(xd.grad, g[f'{name}/gx']) < TOL[prec] and rel_err(pd.grad, g[f'{name}/gpar']) < TOL[prec]


def test_ragged_and_empty_batches():
    spec, n, x, par = cases.transformer_cases(torch.float32)['spline_k8']
    mod = to_module(spec).to(DEV)
    for B in (0, 1, 7, 24):
        y, ld = mod(x[:B].to(DEV), par[:B].to(DEV))
        assert y.shape == (B, n) and ld.shape == (B,)
        if B == 0:
            continue                     # the reference cannot reshape an empty parameter tensor
        y_o, ld_o = cases.double_reference(spec, x[:B], par[:B])
        assert rel_err(y, y_o) < 1e-5 and rel_err(ld, ld_o) < 1e-5


def test_functional_api():
    from tfep_b200.nn.transformers import (affine_transformer, affine_transformer_inverse, moebius_transformer,
                                           volume_preserving_shift_transformer,
                                           volume_preserving_shift_transformer_inverse)
    x, s, a = (cases.normal((9, 5), k).to(DEV) for k in (1, 2, 3))
    y, ld = affine_transformer(x, s, a)
    assert rel_err(y, x.cpu() * torch.exp(a.cpu()) + s.cpu()) < 1e-6 and rel_err(ld, a.cpu().sum(1)) < 1e-6
    xi, _ = affine_transformer_inverse(y, s, a)
    assert rel_err(xi, x) < 1e-5
    ys, lds = volume_preserving_shift_transformer(x, s)
    xs, _ = volume_preserving_shift_transformer_inverse(ys, s)
    assert rel_err(ys, x.cpu() + s.cpu()) < 1e-6 and float(lds.abs().max()) == 0.0 and rel_err(xs, x) < 1e-6
    xv, wv = cases.normal((4, 3, 3), 5), cases.normal((4, 3, 3), 6)
    y, ld = moebius_transformer(xv.to(DEV), wv.to(DEV))
    y_o, ld_o = fo.Moebius(3).forward(xv.reshape(4, 9), wv.reshape(4, 9))
    assert rel_err(y.reshape(4, 9), y_o) < 1e-5 and rel_err(ld, ld_o) < 1e-5
