"""Transformer fused into the output-layer product of the tensor-core conditioner (tfepb_tc_tx; precision='bf16' for
affine / SOS / Moebius / neural-spline flows, forward and backward).

The fused and the separate paths run the SAME products on the same bf16 operand images with fp32 accumulation, and the
epilogue calls the same device math as the transformer kernels, so they must agree to summation order (the log-det is
accumulated with atomics in the fused path): 1e-5 relative.  Against the exact fp32 path the stated bf16 tolerance of
tests/test_gpu_fused.py applies.  Covered: ragged batches (rows beyond the last full tile of 128), feature counts that
leave ragged 16-column chunks, ascending and descending degree orders, both Moebius variants, gradients w.r.t. x and every
parameter (weight-norm g / v and biases of all layers).
"""

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _flow(kind, n_features, order, hidden=2, seed=0):
    from tfep_b200.nn.conditioners.made import generate_degrees
    from tfep_b200.nn.flows import MAF
    from tfep_b200.nn.transformers import AffineTransformer, MoebiusTransformer, NeuralSplineTransformer, SOSPolynomialTransformer
    torch.manual_seed(seed)
    if kind in ('affine_cond', 'spline_cond'):
        # three conditioning features (degree -1) in front: read by the conditioner, copied through by the flow
        deg = torch.cat([torch.full((3,), -1), generate_degrees(n_features - 3, order=order)])
        tr = AffineTransformer() if kind == 'affine_cond' else \
            NeuralSplineTransformer(-torch.ones(n_features - 3) * 2.0, torch.ones(n_features - 3) * 2.5, 8)
        return MAF(deg, tr, hidden_layers=hidden, initialize_identity=False).to(DEV)
    if kind == 'spline_embedded':
        # the reference's MixedMAFMap in full: the periodic features enter the conditioner as (cos, sin) through a
        # PeriodicEmbedding, circular splines map them, open splines the rest
        from tfep_b200.nn.embeddings import PeriodicEmbedding
        from tfep_b200.nn.transformers import MixedTransformer
        ia = [i for i in range(n_features) if i % 2 == 0]
        ib = [i for i in range(n_features) if i % 2 == 1]
        pi = 3.141592653589793
        ta = NeuralSplineTransformer(-torch.ones(len(ia)) * pi, torch.ones(len(ia)) * pi, 8, circular=True)
        tb = NeuralSplineTransformer(-torch.ones(len(ib)) * 2.0, torch.ones(len(ib)) * 2.5, 8)
        emb = PeriodicEmbedding(n_features, [-pi, pi], periodic_indices=ia)
        return MAF(generate_degrees(n_features, order=order), MixedTransformer([ta, tb], [ia, ib]), hidden_layers=hidden,
                   embedding=emb, initialize_identity=False).to(DEV)
    if kind == 'spline_mixed':
        # the reference's MixedMAFMap shape: circular splines on some features, open splines (other options, another domain)
        # on the rest, interleaved
        from tfep_b200.nn.transformers import MixedTransformer
        ia = [i for i in range(n_features) if i % 3 == 0]
        ib = [i for i in range(n_features) if i % 3 != 0]
        ta = NeuralSplineTransformer(-torch.ones(len(ia)) * 3.0, torch.ones(len(ia)) * 3.0, 8, circular=True)
        tb = NeuralSplineTransformer(-torch.ones(len(ib)) * 2.0, torch.ones(len(ib)) * 2.5, 8, identity_boundary_slopes=True)
        return MAF(generate_degrees(n_features, order=order), MixedTransformer([ta, tb], [ia, ib]), hidden_layers=hidden,
                   initialize_identity=False).to(DEV)
    if kind.startswith('spline'):
        lo, hi = -torch.ones(n_features) * 2.0, torch.ones(n_features) * 2.5
        opts = {'spline': dict(circular=True), 'spline_open': dict(), 'spline_id': dict(identity_boundary_slopes=True),
                'spline_learn': dict(learn_lower_bound=True, learn_upper_bound=True),
                'spline_circ_id': dict(circular=True, identity_boundary_slopes=True)}[kind]
        return MAF(generate_degrees(n_features, order=order), NeuralSplineTransformer(lo, hi, 8, **opts), hidden_layers=hidden,
                   initialize_identity=False).to(DEV)
    if kind == 'affine':
        return MAF(generate_degrees(n_features, order=order), AffineTransformer(), hidden_layers=hidden,
                   initialize_identity=False).to(DEV)
    if kind == 'sos':
        return MAF(generate_degrees(n_features, order=order), SOSPolynomialTransformer(2), hidden_layers=hidden,
                   initialize_identity=False).to(DEV)
    unit = kind == 'moebius_unit'
    return MAF(generate_degrees(n_features, order=order, repeats=3), MoebiusTransformer(dimension=3, unit_sphere=unit),
               hidden_layers=hidden, initialize_identity=False).to(DEV)


def _input(kind, batch, n_features, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, n_features, generator=g)
    if kind == 'moebius_unit':
        x = x.view(batch, -1, 3)
        x = (x / x.norm(dim=-1, keepdim=True)).reshape(batch, n_features)
    return x.to(DEV)


def _run(maf, x, cy, cl, precision, fuse):
    maf.precision = precision
    maf.fuse_transformer = fuse
    maf.zero_grad(set_to_none=True)
    xg = x.clone().requires_grad_(True)
    y, ld = maf(xg)
    loss = (y * cy).sum()
    if ld.requires_grad:
        loss = loss + (ld * cl).sum()
    loss.backward()
    return y.detach(), ld.detach(), xg.grad.clone(), {k: p.grad.clone() for k, p in maf.named_parameters()}


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


CASES = [('spline', 23, 'ascending', 300), ('spline_open', 9, 'descending', 200), ('spline_id', 12, 'ascending', 129),
         ('spline_learn', 10, 'descending', 260), ('spline_circ_id', 8, 'ascending', 64), ('spline', 66, 'ascending', 1000),
         ('spline_mixed', 20, 'ascending', 300), ('spline_mixed', 66, 'descending', 500), ('spline_embedded', 24, 'ascending', 300),
         ('affine_cond', 20, 'ascending', 300), ('spline_cond', 13, 'descending', 200),
         ('affine', 37, 'ascending', 300), ('affine', 64, 'descending', 128), ('sos', 40, 'ascending', 300),
         ('sos', 9, 'descending', 77), ('sos', 300, 'ascending', 1000), ('moebius', 36, 'ascending', 300),
         ('moebius', 33, 'descending', 513), ('moebius_unit', 30, 'ascending', 200), ('moebius', 300, 'descending', 700)]


@pytest.mark.parametrize('kind,n_features,order,batch', CASES)
def test_fused_epilogue_matches_the_separate_kernels(kind, n_features, order, batch):
    maf = _flow(kind, n_features, order)
    assert maf._tc_tx_plan() is not None, maf._tctx_why
    x = _input(kind, batch, n_features, 5)
    g = torch.Generator().manual_seed(6)
    cy, cl = torch.randn(batch, n_features, generator=g).to(DEV), torch.randn(batch, generator=g).to(DEV)
    y0, ld0, gx0, gp0 = _run(maf, x, cy, cl, 'bf16', fuse=False)
    y1, ld1, gx1, gp1 = _run(maf, x, cy, cl, 'bf16', fuse=True)
    assert _rel(y1, y0) < 1e-5 and float((ld1 - ld0).abs().max()) < 1e-4 * (1 + float(ld0.abs().max()))
    # gradients: the cotangent of the parameters reaches the products below as the same bf16 image in both paths
    # (the spline VJP has long cancellation-prone expressions: its fast-math evaluation in the epilogue differs from the exact
    # kernel by ~1e-6 relative per cotangent, which flips bf16 roundings of the operand image)
    tol = 1e-3 if kind.startswith('spline') else 1e-4
    assert _rel(gx1, gx0) < tol, _rel(gx1, gx0)
    for k in gp0:
        assert _rel(gp1[k], gp0[k]) < tol, (k, _rel(gp1[k], gp0[k]))
    # and against the exact path: bf16 operand rounding
    y32, ld32, gx32, gp32 = _run(maf, x, cy, cl, 'fp32', fuse=False)
    # (the log-det is a sum over the features: the tolerance scales with its size)
    assert _rel(y1, y32) < 2e-2 and float((ld1 - ld32).abs().mean()) < 5e-2 + 5e-4 * float(ld32.abs().mean())
    assert _rel(gx1, gx32) < 8e-2
    for k in gp32:
        assert _rel(gp1[k], gp32[k]) < 8e-2, k


def test_fused_epilogue_inference_and_eligibility():
    """No autograd: nothing is saved, same results; flows the epilogue does not cover keep the separate kernels."""
    from tfep_b200.nn.conditioners.made import generate_degrees
    from tfep_b200.nn.flows import MAF
    from tfep_b200.nn.transformers import SOSPolynomialTransformer, SymmetrizedMoebiusTransformer
    maf = _flow('sos', 50, 'ascending')
    x = _input('sos', 1000, 50, 1)
    maf.precision = 'bf16'
    with torch.no_grad():
        y1, ld1 = maf(x)
        maf.fuse_transformer = False
        y0, ld0 = maf(x)
    assert _rel(y1, y0) < 1e-5 and float((ld1 - ld0).abs().max()) < 1e-4 * (1 + float(ld0.abs().max()))
    sos3 = MAF(generate_degrees(8), SOSPolynomialTransformer(3), initialize_identity=False, precision='bf16').to(DEV)
    assert sos3._tc_tx_plan() is None and 'polynomials' in sos3._tctx_why
    sym = MAF(generate_degrees(12, repeats=3), SymmetrizedMoebiusTransformer(dimension=3), initialize_identity=False,
              precision='bf16').to(DEV)
    assert sym._tc_tx_plan() is None
    with torch.no_grad():
        for m, n in ((sos3, 8), (sym, 12)):
            y, ld = m(torch.randn(64, n, device=DEV))
            assert torch.isfinite(y).all() and torch.isfinite(ld).all()


def test_fused_epilogue_in_a_training_step():
    """Three stacked layers (SOS, Moebius, affine), KL-style loss, one optimizer step: the loss and the updated parameters
    agree with the separate-kernel path."""
    from tfep_b200.nn.flows import SequentialFlow

    def build():
        return SequentialFlow(_flow('sos', 30, 'ascending', seed=1), _flow('moebius', 30, 'descending', seed=2),
                              _flow('affine', 30, 'ascending', seed=3)).to(DEV)

    x = _input('sos', 600, 30, 9)
    out = []
    for fuse in (False, True):
        flow = build()
        for m in flow:
            m.precision = 'bf16'
            m.fuse_transformer = fuse
        opt = torch.optim.SGD(flow.parameters(), lr=1e-2)
        y, ld = flow(x)
        loss = (0.5 * (y * y).sum(dim=1) - ld).mean()
        loss.backward()
        opt.step()
        out.append((float(loss), torch.cat([p.detach().flatten() for p in flow.parameters()])))
    assert abs(out[0][0] - out[1][0]) < 1e-4 * (1 + abs(out[0][0]))
    assert _rel(out[1][1], out[0][1]) < 1e-5


@pytest.mark.parametrize('kind,n_features', [('sos', 21), ('moebius', 27), ('affine', 10)])
@pytest.mark.parametrize('batch', [1, 3, 129])
@pytest.mark.parametrize('which', ['x', 'parameters'])
def test_fused_epilogue_partial_gradients_and_tiny_batches(kind, n_features, batch, which):
    """Only x, or only the parameters, require a gradient (frozen conditioner / plain training step), on batches smaller than
    one reduction block of the weight gradient: same results as the separate kernels."""
    maf = _flow(kind, n_features, 'ascending', seed=11)
    maf.precision = 'bf16'
    for p in maf.parameters():
        p.requires_grad_(which == 'parameters')
    x = _input(kind, batch, n_features, 12)
    g = torch.Generator().manual_seed(13)
    cy = torch.randn(batch, n_features, generator=g).to(DEV)
    out = []
    for fuse in (False, True):
        maf.fuse_transformer = fuse
        maf.zero_grad(set_to_none=True)
        xg = x.clone().requires_grad_(which == 'x')
        y, ld = maf(xg)
        ((y * cy).sum() + (ld.sum() if ld.requires_grad else 0.0)).backward()
        grads = [xg.grad] if which == 'x' else [p.grad for p in maf.parameters()]
        assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
        out.append((y.detach(), ld.detach(), torch.cat([gr.flatten() for gr in grads])))
    assert _rel(out[1][0], out[0][0]) < 1e-5
    assert float((out[1][1] - out[0][1]).abs().max()) < 1e-4 * (1 + float(out[0][1].abs().max()))
    assert _rel(out[1][2], out[0][2]) < 2e-4
