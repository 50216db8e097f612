"""MAF forward / inverse / gradients on the GPU against the reference's golden vectors and the oracle."""

import math

import numpy as np
import pytest
import torch

from helpers import cfg_flow_modules, golden, rel_err, to_maf, to_wrapper
from oracle import cases
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = {'f32': 1e-5, 'f64': 1e-10}
DT = {'f32': torch.float32, 'f64': torch.float64}


@pytest.fixture(params=['f32', 'f64'])
def prec(request):
    old = torch.get_default_dtype()
    torch.set_default_dtype(DT[request.param])
    yield request.param
    torch.set_default_dtype(old)


def test_forward_and_inverse_against_golden(prec):
    g = golden(f'maf_{prec}.npz')
    g64 = golden('maf_f64.npz')
    for name, case in cases.maf_cases(DT[prec]).items():
        _, sd = cases.build_oracle(case, DT[prec])
        maf = to_maf(case, sd, DEV, DT[prec])
        with torch.no_grad():
            y, ld = maf(case['x'].to(DEV))
        assert rel_err(y, g[f'{name}/y']) < 2 * TOL[prec], name
        assert rel_err(ld, g[f'{name}/ld']) < 2 * TOL[prec], name
        if not case['invertible']:
            with pytest.raises(NotImplementedError):
                maf.inverse(y)
            continue
        with torch.no_grad():
            xi, ldi = maf.inverse(torch.from_numpy(g[f'{name}/y']).to(DEV))
        # The inverse divides by dy/dx, so its error is the forward tolerance times the conditioning of the case.  The
        # yardstick for that is the reference itself: how far its fp32 inverse lies from its fp64 inverse (golden files of
        # both precisions share the seeds).  fp32: within 2e-5, or 4 x the reference's own fp32 error where that is larger.
        tol_x = tol_ld = 2 * TOL[prec]
        if prec == 'f32':
            tol_x = max(tol_x, 4 * rel_err(g[f'{name}/xinv'], g64[f'{name}/xinv']))
            tol_ld = max(tol_ld, 4 * rel_err(g[f'{name}/ldinv'], g64[f'{name}/ldinv']))
            assert tol_x < 20 * TOL[prec] and tol_ld < 20 * TOL[prec], name          # never looser than before
        assert rel_err(xi, g[f'{name}/xinv']) < tol_x, (name, tol_x)
        assert rel_err(ldi, g[f'{name}/ldinv']) < tol_ld, (name, tol_ld)


def test_generic_autoregressive_flow_matches_packed_path():
    """AutoregressiveFlow.forward/inverse (reference layout, one conditioner pass per degree group) and
    the packed MAF path (single sweep) are the same map."""
    from tfep_b200.nn.flows.autoregressive import AutoregressiveFlow
    for name, case in cases.maf_cases(torch.float64).items():
        _, sd = cases.build_oracle(case, torch.float64)
        maf = to_maf(case, sd, DEV, torch.float64)
        x = case['x'].to(DEV)
        with torch.no_grad():
            y, ld = maf(x)
            y2, ld2 = AutoregressiveFlow.forward(maf, x)
            assert rel_err(y2, y) < 1e-12 and rel_err(ld2, ld) < 1e-12, name
            if case['invertible']:
                xi, ldi = maf.inverse(y)
                xi2, ldi2 = AutoregressiveFlow.inverse(maf, y)
                assert rel_err(xi2, xi) < 1e-10 and rel_err(ldi2, ldi) < 1e-10, name


def test_persistent_sweep_matches_host_driven_sweep(prec):
    """tfepb_maf_inverse_sweep (one persistent kernel, tile state in shared memory) against the same
    degree-ordered sweep driven from the host with the stand-alone GEMM / transformer kernels, on every
    invertible MAF configuration, for batches that are not multiples of the 64 / 32-sample tile."""
    tol = 20 * TOL[prec]
    for name, case in cases.maf_cases(DT[prec]).items():
        if not case['invertible']:
            continue
        _, sd = cases.build_oracle(case, DT[prec])
        maf = to_maf(case, sd, DEV, DT[prec])
        x = case['x'].to(DEV)
        reps = -(-150 // x.shape[0])
        xx = torch.cat([x] * reps)[:150] if x.shape[0] < 150 else x
        with torch.no_grad():
            y, _ = maf(xx)
            for b in (1, 33, 150):
                xi, ldi = maf.inverse(y[:b])
                if case.get('embedding') is None:
                    xh, ldh = maf._inverse_host_sweep(y[:b])
                else:       # the host-driven sweep has no embedding: compare with the generic one-pass-per-degree inverse
                    from tfep_b200.nn.flows.autoregressive import AutoregressiveFlow
                    xh, ldh = AutoregressiveFlow.inverse(maf, y[:b])
                assert rel_err(xi, xh) < tol and rel_err(ldi, ldh) < tol, (name, b)
                xi2, ldi2 = maf.inverse(y[:b])
                assert torch.equal(xi, xi2) and torch.equal(ldi, ldi2), name       # deterministic


def test_round_trip_and_conditioning_untouched(prec):
    """inverse(forward(x)) == x, log-dets cancel, conditioning features pass through
    (reference tests/nn/flows/test_maf.py:226-295)."""
    for name, case in cases.maf_cases(DT[prec]).items():
        if not case['invertible']:
            continue
        _, sd = cases.build_oracle(case, DT[prec])
        maf = to_maf(case, sd, DEV, DT[prec])
        x = case['x'].to(DEV)
        with torch.no_grad():
            y, ld = maf(x)
            xi, ldi = maf.inverse(y)
        fixed = (case['degrees_in'] == -1).nonzero().flatten().to(DEV)
        assert torch.equal(y[:, fixed], x[:, fixed]), name
        spec = case['spec']
        if isinstance(spec, (fo.Spline, fo.Mixed)) or (isinstance(spec, fo.Shift) and spec.periodic_indices is not None):
            continue                                     # periodic features come back modulo the period (and the
                                                         # reference's shift wrap `v % P + lower` is not its own inverse)
        assert rel_err(xi, x) < 100 * TOL[prec], name
        assert rel_err(ld + ldi, torch.zeros_like(ld)) < 100 * TOL[prec], name


def test_gradients_against_reference_autograd(prec):
    """d loss / d x and d loss / d (g, v, bias) vs PyTorch autograd through the real reference."""
    g = golden(f'maf_{prec}.npz')
    dtype = DT[prec]
    for name, case in cases.maf_cases(dtype).items():
        _, sd = cases.build_oracle(case, dtype)
        maf = to_maf(case, sd, DEV, dtype)
        x = case['x'].to(DEV).requires_grad_(True)
        y, ld = maf(x)
        cy, cl = cases.normal(tuple(y.shape), 78, dtype).to(DEV), cases.normal(tuple(ld.shape), 79, dtype).to(DEV)
        loss = (y * cy).sum() + ((ld * cl).sum() if ld.requires_grad else 0.0)
        loss.backward()
        tol = 30 * TOL[prec]
        assert rel_err(x.grad, g[f'{name}/gx']) < tol, name
        for k, p in maf.named_parameters():
            ref = g[f'{name}/grad/{k}']
            scale = 1 + float(np.abs(ref).max())
            assert float((p.grad.cpu().double() - torch.from_numpy(ref).double()).abs().max()) / scale < tol, (name, k)


def test_autoregressive_property(prec):
    """y_i depends on x_j only if degree(j) < degree(i), or j == i (reference tests/nn/__init__.py:25-96)."""
    for name in ('spline_desc_cond', 'moebius_d2_cond', 'repeated_degrees', 'mixed_splines'):
        case = cases.maf_cases(DT[prec])[name]
        _, sd = cases.build_oracle(case, DT[prec])
        maf = to_maf(case, sd, DEV, DT[prec])
        x = case['x'].to(DEV).requires_grad_(True)
        y, _ = maf(x)
        deg = case['degrees_in']
        dim = case['spec'].dimension if isinstance(case['spec'], fo.Moebius) else 1
        for i in range(y.shape[1]):
            gi, = torch.autograd.grad(y[:, i].sum(), x, retain_graph=True)
            dep = (gi.abs().sum(0) > 0).cpu()
            for j in range(y.shape[1]):
                same_block = dim > 1 and deg[i] == deg[j] and deg[i] != -1 and abs(i - j) < dim
                allowed = (deg[j] < deg[i] and deg[i] != -1) or j == i or same_block
                assert allowed or not bool(dep[j]), (name, i, j)


def test_identity_initialisation_is_the_identity_map():
    from tfep_b200.nn.conditioners import generate_degrees
    from tfep_b200.nn.flows import MAF
    from tfep_b200.nn.transformers import (AffineTransformer, MoebiusTransformer, NeuralSplineTransformer,
                                           SOSPolynomialTransformer)
    x = cases.uniform((32, 6), 3, -0.9, 0.9).to(DEV)
    for t, rep in ((AffineTransformer(), 1), (SOSPolynomialTransformer(2), 1), (SOSPolynomialTransformer(3), 1),
                   (MoebiusTransformer(2), 2),
                   (NeuralSplineTransformer(torch.full((6,), -1.), torch.full((6,), 1.), 5), 1),
                   (NeuralSplineTransformer(torch.full((6,), -1.), torch.full((6,), 1.), 5, circular=True), 1)):
        for order in ('ascending', 'descending'):
            maf = MAF(generate_degrees(6, order=order, repeats=rep), transformer=t).to(DEV)
            with torch.no_grad():
                y, ld = maf(x)
            assert rel_err(y, x) < 1e-5 and float(ld.abs().max()) < 1e-4, (type(t).__name__, order)


def test_config_slices_fp32_parity_and_fp64_accuracy():
    """BASELINE.json configurations (reduced batch; cfg3 / cfg5 at reduced D): the fp32 path is within 1e-5
    (relative) of the reference's fp32 CPU result; the error against the fp64 reference is reported too."""
    g = golden('cfg_slices.npz')
    for cfg, nl, B, D in (('cfg1', 2, 64, None), ('cfg2', 4, 64, None), ('cfg3', 6, 32, 30), ('cfg5', 2, 32, 24)):
        seq, _ = cfg_flow_modules(cfg, DEV, n_layers=nl, D=D)
        x = torch.from_numpy(g[f'{cfg}/x']).to(DEV)
        with torch.no_grad():
            y, ld = seq(x)
        assert rel_err(y, g[f'{cfg}/f32/y']) < 1e-5, cfg
        assert rel_err(ld, g[f'{cfg}/f32/ld']) < 1e-5 * nl, cfg
        assert rel_err(y, g[f'{cfg}/f64/y']) < 1e-4 and rel_err(ld, g[f'{cfg}/f64/ld']) < 1e-4, cfg
        if cfg != 'cfg3':
            with torch.no_grad():
                xi, ldi = seq.inverse(torch.from_numpy(g[f'{cfg}/f32/y']).to(DEV))
            assert rel_err(ldi, g[f'{cfg}/f32/ldinv']) < 2e-4, cfg
            if cfg != 'cfg2':
                assert rel_err(xi, g[f'{cfg}/f32/xinv']) < 2e-4, cfg
            else:
                d = (xi.cpu() - torch.from_numpy(g[f'{cfg}/f32/xinv'])).abs()
                assert float(torch.minimum(d, (2 * math.pi - d).abs()).max()) < 2e-4


def test_headline_config_full_size_properties():
    """cfg2 at the BASELINE.json batch (65536): size-independent properties instead of an oracle run --
    determinism, batch-slicing invariance, round trip modulo the period and log-det cancellation."""
    seq, _ = cfg_flow_modules('cfg2', DEV)
    x = cases.cfg_input('cfg2', 65536).to(DEV)
    with torch.no_grad():
        y, ld = seq(x)
        y2, ld2 = seq(x)
        assert torch.equal(y, y2) and torch.equal(ld, ld2)
        ys, lds = seq(x[1000:1777])
        assert torch.equal(ys, y[1000:1777]) and torch.equal(lds, ld[1000:1777])
        assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(ld).all())
        assert float(y.min()) >= -math.pi - 1e-5 and float(y.max()) <= math.pi + 1e-5
        sub = slice(0, 4096)
        xi, ldi = seq.inverse(y[sub])
        d = (xi - x[sub]).abs()
        assert float(torch.minimum(d, (2 * math.pi - d).abs()).max()) < 2e-3
        assert float((ld[sub] + ldi).abs().max()) < 5e-3


def test_host_pipeline_matches_direct_call():
    """Chunked, stream-overlapped evaluation from / to pinned host memory == one direct call."""
    from tfep_b200.utils.host_pipeline import HostPipeline
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=2)
    x = cases.cfg_input('cfg2', 5000).pin_memory()
    with torch.no_grad():
        y, ld = seq(x.to(DEV))
    pipe = HostPipeline(seq, 5000, 66, DEV, n_chunks=3)
    for _ in range(2):
        yh, ldh = pipe(x)
        torch.cuda.synchronize()
        assert torch.equal(yh, y.cpu()) and torch.equal(ldh, ld.cpu())
    # overlapped across batches (wait=False) and CUDA-graph mode: every staging set must hold the same result
    for maf in seq:
        maf.precision = 'bf16'
    with torch.no_grad():
        yb, ldb = seq(x.to(DEV))
    for mode in ('streams', 'graph'):
        pipe = HostPipeline(seq, 5000, 66, DEV, n_chunks=1, depth=3)
        outs = [pipe(x, wait=False) if mode == 'streams' else pipe.step_graph(x) for _ in range(7)]
        pipe.join()
        torch.cuda.synchronize()
        for yh, ldh in outs[-3:]:
            assert torch.equal(yh, yb.cpu()) and torch.equal(ldh, ldb.cpu()), mode


def test_wrapper_flows_around_the_maf_kernels(prec):
    """PartialFlow / CenteredCentroidFlow / OrientedFlow (reference nn/flows/{partial,centroid,oriented}.py) around the
    CUDA MAF: forward, inverse and the gradient w.r.t. x (autograd through the frame algebra AND the hand-written
    MAF backward kernels) against the reference's golden vectors."""
    g = golden(f'wrappers_{prec}.npz')
    for name, case in cases.wrapper_cases(DT[prec]).items():
        _, sd = cases.build_oracle(case['inner'], DT[prec])
        flow = to_wrapper(case, to_maf(case['inner'], sd, DEV, DT[prec]), DT[prec]).to(DEV)
        x = case['x'].to(DEV).requires_grad_(True)
        y, ld = flow(x)
        assert rel_err(y, g[f'{name}/y']) < 5 * TOL[prec] and rel_err(ld, g[f'{name}/ld']) < 5 * TOL[prec], name
        cy, cl = cases.normal(tuple(y.shape), 78, DT[prec]).to(DEV), cases.normal(tuple(ld.shape), 79, DT[prec]).to(DEV)
        ((y * cy).sum() + (ld * cl).sum()).backward()
        assert rel_err(x.grad, g[f'{name}/gx']) < 50 * TOL[prec], name
        if case['invertible']:
            with torch.no_grad():
                xi, ldi = flow.inverse(torch.from_numpy(g[f'{name}/y']).to(DEV))
            assert rel_err(xi, g[f'{name}/xinv']) < 50 * TOL[prec] and rel_err(ldi, g[f'{name}/ldinv']) < 50 * TOL[prec], name


def test_mixed_embedding_in_front_of_the_conditioner():
    """MixedEmbedding of a PeriodicEmbedding (CUDA kernel) and a learnable FlipInvariantEmbedding (reference
    mafembed.py:174-446) against the reference's golden output, and as the `embedding=` of a MAF (conditioner input
    degrees from get_degrees_out; forward / inverse round trip through the generic per-degree inverse)."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_api import _load_embedding
    from tfep_b200.nn.flows import MAF
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        emb, g, deg = _load_embedding('mixed_periodic', DEV)
        x = torch.from_numpy(g['mixed_periodic/x']).to(DEV).requires_grad_(True)
        y = emb(x)
        assert rel_err(y, g['mixed_periodic/y']) < 1e-5
        (y * cases.normal(tuple(y.shape), 42).to(DEV)).sum().backward()
        assert rel_err(x.grad, g['mixed_periodic/gx']) < 1e-4
        assert torch.equal(emb.get_degrees_out(deg).cpu(), torch.from_numpy(g['mixed_periodic/deg']))
        torch.manual_seed(3)
        maf = MAF(degrees_in=deg, embedding=emb, initialize_identity=False).to(DEV)
        xs = cases.normal((50, 9), 43).to(DEV)
        with torch.no_grad():
            ys, ld = maf(xs)
            xi, ldi = maf.inverse(ys)
        assert rel_err(xi, xs) < 1e-4 and rel_err(ld + ldi, torch.zeros_like(ld)) < 1e-4
    finally:
        torch.set_default_dtype(old)


def test_gradients_through_the_inverse(prec):
    """MAF.inverse under autograd (implicit differentiation of F(x) = y on the hand-written backward kernels) against
    autograd through the oracle's n_degrees-pass inverse loop (= what the reference differentiates): gradients
    w.r.t. y and w.r.t. the biases of every conditioner layer, for elementwise, vector-block (Moebius) and mixed
    transformers, with and without conditioning features."""
    dtype = DT[prec]
    tol = {'f32': 2e-3, 'f64': 1e-8}[prec]
    for name in ('affine_asc', 'affine_cond_h1', 'spline_desc_cond', 'moebius_d3', 'moebius_d2_cond', 'mixed_splines',
                 'repeated_degrees', 'symmoebius_d3', 'spline_embed_periodic'):
        case = cases.maf_cases(dtype)[name]
        oracle, sd = cases.build_oracle(case, dtype)
        maf = to_maf(case, sd, DEV, dtype)
        with torch.no_grad():
            y0, _ = maf(case['x'].to(DEV))
        cx, cl = cases.normal(tuple(y0.shape), 61, dtype), cases.normal((y0.shape[0],), 62, dtype)
        # oracle: autograd through the loop, leaves = y and the (effective weight, bias) pairs
        yo = y0.cpu().clone().requires_grad_(True)
        oracle.layers = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in oracle.layers]
        xo, ldo = oracle.inverse(yo)
        ((xo * cx).sum() + (ldo * cl).sum()).backward()
        # tfep_b200
        yg = y0.clone().requires_grad_(True)
        xi, ldi = maf.inverse(yg)
        ((xi * cx.to(DEV)).sum() + (ldi * cl.to(DEV)).sum()).backward()
        assert rel_err(xi.detach(), xo.detach()) < tol and rel_err(ldi.detach(), ldo.detach()) < tol, name
        scale = 1 + float(yo.grad.abs().max())
        assert float((yg.grad.cpu() - yo.grad).abs().max()) / scale < tol, name
        lin = maf._conditioner._linear_layers()
        for (w_o, b_o), layer in zip(oracle.layers, lin):
            gb = layer.bias.grad
            assert gb is not None, name
            assert float((gb.cpu() - b_o.grad).abs().max()) / (1 + float(b_o.grad.abs().max())) < tol, name


def test_wrapper_flows_fused_kernels(prec):
    """Outside autograd the wrapper flows run their fused pre / post kernels (tfep_b200/csrc/frames.cu): same results as
    the reference's golden vectors and as the differentiable tensor-algebra path, forward and inverse."""
    g = golden(f'wrappers_{prec}.npz')
    for name, case in cases.wrapper_cases(DT[prec]).items():
        if all(kind == 'partial' for kind, _ in case['layers']):
            continue
        _, sd = cases.build_oracle(case['inner'], DT[prec])
        flow = to_wrapper(case, to_maf(case['inner'], sd, DEV, DT[prec]), DT[prec]).to(DEV)
        x = case['x'].to(DEV)
        with torch.no_grad():
            y, ld = flow(x)                                           # kernels
        ya, lda = flow(x.clone().requires_grad_(True))                # tensor algebra
        assert flow._frame is not None, name
        assert rel_err(y, g[f'{name}/y']) < 5 * TOL[prec] and rel_err(ld, g[f'{name}/ld']) < 5 * TOL[prec], name
        assert rel_err(y, ya.detach()) < 5 * TOL[prec] and rel_err(ld, lda.detach()) < 5 * TOL[prec], name
        if case['invertible']:
            with torch.no_grad():
                xi, ldi = flow.inverse(torch.from_numpy(g[f'{name}/y']).to(DEV))
            assert rel_err(xi, g[f'{name}/xinv']) < 50 * TOL[prec] and rel_err(ldi, g[f'{name}/ldinv']) < 50 * TOL[prec], name
    # a point exactly on the axis (parallel case of the frame) and ragged batch sizes
    case = cases.wrapper_cases(DT[prec])['oriented']
    _, sd = cases.build_oracle(case['inner'], DT[prec])
    flow = to_wrapper(case, to_maf(case['inner'], sd, DEV, DT[prec]), DT[prec]).to(DEV)
    x = case['x'].to(DEV).clone()
    x[0, 0:3] = torch.tensor([2.5, 0.0, 0.0])
    x[1, 0:3] = torch.tensor([-1.5, 0.0, 0.0])
    for B in (1, 2, 7):
        with torch.no_grad():
            y, ld = flow(x[:B])
        ya, lda = flow(x[:B].clone().requires_grad_(True))
        assert rel_err(y, ya.detach()) < 5 * TOL[prec] and rel_err(ld, lda.detach()) < 5 * TOL[prec], B


def test_empty_batches(prec):
    """B = 0 through every MAF path: exact forward / inverse sweep for each transformer family, embedded and
    conditioned flows, and the general tensor-core conditioner (precision='bf16' outside the fused kernel)."""
    for name, case in cases.maf_cases(DT[prec]).items():
        _, sd = cases.build_oracle(case, DT[prec])
        maf = to_maf(case, sd, DEV, DT[prec])
        n = case['x'].shape[1]
        x = torch.empty(0, n, dtype=DT[prec], device=DEV)
        with torch.no_grad():
            y, ld = maf(x)
            assert y.shape == (0, n) and ld.shape == (0,), name
            if case['invertible']:
                xi, ldi = maf.inverse(x)
                assert xi.shape == (0, n) and ldi.shape == (0,), name
            if prec == 'f32':
                maf.precision = 'bf16'
                y, ld = maf(x)
                assert y.shape == (0, n) and ld.shape == (0,), name
