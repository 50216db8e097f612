"""Masked linear layers and the MADE conditioner on the GPU against the oracle."""

import pytest
import torch

from helpers import rel_err, to_maf
from oracle import cases
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.float64, 1e-12)], ids=['f32', 'f64'])
@pytest.mark.parametrize('shape', [(1, 3, 5), (130, 66, 93), (257, 93, 132), (64, 328, 70)])
def test_masked_linear_forward_backward(dtype, tol, shape):
    from tfep_b200.nn.masked import masked_linear
    B, K, N = shape
    x, w, b = cases.normal((B, K), 1, dtype), cases.normal((N, K), 2, dtype), cases.normal((N,), 3, dtype)
    mask = (cases.uniform((N, K), 4, 0, 1, dtype) > 0.5).to(dtype)
    gy = cases.normal((B, N), 5, dtype)
    ref_in = [t.clone().requires_grad_(True) for t in (x, w, b)]
    y_o = torch.nn.functional.linear(ref_in[0], ref_in[1] * mask, ref_in[2])
    (y_o * gy).sum().backward()
    dev_in = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    y = masked_linear(dev_in[0], dev_in[1], dev_in[2], mask.to(DEV))
    (y * gy.to(DEV)).sum().backward()
    scale = float(K) ** 0.5
    assert rel_err(y, y_o) < tol * scale
    for a, o in zip(dev_in, ref_in):
        assert rel_err(a.grad, o.grad) < tol * max(scale, float(B) ** 0.5), shape


def test_masked_linear_module_and_weight_norm():
    from tfep_b200.nn import masked
    torch.manual_seed(0)
    mask = torch.tril(torch.ones(5, 8))
    lin = masked.masked_weight_norm(masked.MaskedLinear(8, 5, mask=mask)).to(DEV)
    x = cases.normal((20, 8), 1).to(DEV)
    y = lin(x)
    w = fo.effective_weight(lin.weight_v.detach().cpu(), lin.weight_g.detach().cpu(), mask)
    assert rel_err(y, torch.nn.functional.linear(x.cpu(), w, lin.bias.detach().cpu())) < 1e-5
    y.sum().backward()
    assert float(lin.weight_v.grad[mask.to(DEV) == 0].abs().max()) == 0.0     # reference test_masked.py:173-218


@pytest.mark.parametrize('name', ['affine_asc', 'affine_desc_nown', 'affine_cond_h1', 'affine_h4', 'spline_circ'])
def test_made_forward_and_gradients(name):
    """MADE.forward (packed, fused ELU) returns the reference layout; gradients w.r.t. x, g, v, bias."""
    case = cases.maf_cases(torch.float64)[name]
    oracle, sd = cases.build_oracle(case, torch.float64)
    made = to_maf(case, sd, DEV, torch.float64)._conditioner
    x = case['x']
    # oracle with autograd through its torch ops
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o2 = fo.MafOracle(case['degrees_in'], case['spec'], case['hidden_layers'], case['weight_norm']).load(sd_g)
    xg = x.clone().requires_grad_(True)
    par_o = o2.parameters_of(xg)
    c = cases.normal(tuple(par_o.shape), 7, torch.float64)
    (par_o * c).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    par = made(xd)
    assert rel_err(par, par_o) < 1e-11
    (par * c.to(DEV)).sum().backward()
    assert rel_err(xd.grad, xg.grad) < 1e-10
    for k, p in made.named_parameters():
        # plain autograd through weight-norm yields 0/0 on fully masked rows; the reference zeroes those
        # entries with gradient hooks (nn/masked.py:400-402), which is what the product returns
        ref = torch.nan_to_num(sd_g['_conditioner.' + k].grad, nan=0.0)
        assert rel_err(p.grad, ref) < 1e-10, k


def test_autoregressive_property_of_conditioner():
    """Output of degree d must not depend on inputs of degree >= d (reference tests/nn/conditioners/test_made.py:122-143)."""
    case = cases.maf_cases(torch.float64)['affine_desc_nown']
    _, sd = cases.build_oracle(case, torch.float64)
    maf = to_maf(case, sd, DEV, torch.float64)
    made = maf._conditioner
    x = case['x'].to(DEV).requires_grad_(True)
    par = made(x)
    deg_in = case['degrees_in']
    deg_out = case['spec'].degrees_out(deg_in)
    for o in range(par.shape[1]):
        g, = torch.autograd.grad(par[:, o].sum(), x, retain_graph=True)
        dep = (g.abs().sum(0) > 0).cpu()
        assert not bool((dep & (deg_in >= deg_out[o])).any())
