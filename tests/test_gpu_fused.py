"""Fused tcgen05 / TMEM / bulk-copy MAF layer (precision='bf16') against the oracle.

Stated tolerance of the bf16 tensor-core variant (north_star: "within a stated tolerance against an fp64
reference"): operands (x, weights, hidden activations) are rounded to bf16 (8-bit mantissa), accumulation
and the whole spline epilogue are fp32.  Against the fp64 reference of one MAF layer of the headline
configuration that gives |dy| <= 2e-2 (mean ~6e-4) modulo the period and |d logdet| <= 5e-2 (mean ~3e-3);
against a reference that applies the SAME bf16 roundings the kernel must agree to 2e-3 (y) / 5e-3 (logdet) (summation order and
fast-math intrinsics only) -- that second check is what pins the GEMM chain, layouts and schedule.
"""

import math

import pytest
import torch

from helpers import cfg_flow_modules
from oracle import cases

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _bf(t):
    return t.to(torch.bfloat16).double()


def _emulated_layer(oracle, x):
    """One MAF layer in double with the kernel's bf16 operand roundings and log2(e) pre-scalings
    (tfep_b200/_fused.py: hidden accumulators hold log2(e) (W a + b), activations travel as log2(e) ELU)."""
    L2E = 1.4426950408889634
    (w1, b1), (w2, b2), (w3, b3) = [(w.double(), b.double()) for w, b in oracle.layers]

    def act(t):                          # log2(e) ELU(t / log2(e))
        return torch.where(t > 0, t, L2E * (torch.exp2(t) - 1))

    a1 = act(_bf(x) @ _bf((w1 * L2E).float()).T + b1 * L2E)
    a2 = act(_bf(a1.float()) @ _bf(w2.float()).T + b2 * L2E)
    # rows feeding softmax / softplus (parameters 0..23 of every feature) are stored pre-multiplied by log2(e)
    n_feat = w3.shape[0] // 25
    scale = torch.ones(w3.shape[0], dtype=torch.double)
    scale[:24 * n_feat] = L2E
    par = (_bf(a2.float()) @ _bf((w3 * (scale / L2E)[:, None]).float()).T) / scale + b3
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        return cases.as_double(oracle.transformer).forward(x.double(), par)
    finally:
        torch.set_default_dtype(old)


def _circ(a, b):
    d = (a.double().cpu() - b.double().cpu()).abs()
    return torch.minimum(d, (2 * math.pi - d).abs())


@pytest.mark.parametrize('batch', [1, 127, 128, 300, 1000])
def test_layers_against_bf16_emulation_and_fp64(batch):
    seq, flows = cfg_flow_modules('cfg2', DEV, n_layers=2)
    x = cases.cfg_input('cfg2', batch)
    for maf, (oracle, _) in zip(seq, flows):        # ascending and descending degree layers
        maf.precision = 'bf16'
        with torch.no_grad():
            y, ld = maf(x.to(DEV))
        y_e, ld_e = _emulated_layer(oracle, x)
        assert float(_circ(y, y_e).max()) < 2e-3 and float((ld.cpu().double() - ld_e).abs().max()) < 5e-3
        y64, ld64 = oracle.forward(x)               # fp32 oracle ~ fp64 at this scale
        assert float(_circ(y, y64).max()) < 2e-2 and float(_circ(y, y64).mean()) < 2e-3
        assert float((ld.cpu() - ld64).abs().max()) < 5e-2 and float((ld.cpu() - ld64).abs().mean()) < 8e-3
        assert y.shape == (batch, 66) and ld.shape == (batch,)


def test_full_batch_properties_and_agreement_with_fp32_path():
    seq, _ = cfg_flow_modules('cfg2', DEV)
    x = cases.cfg_input('cfg2', 65536).to(DEV)
    with torch.no_grad():
        y32, ld32 = seq(x)
        for maf in seq:
            maf.precision = 'bf16'
        y, ld = seq(x)
        y2, ld2 = seq(x)
        assert torch.equal(y, y2) and torch.equal(ld, ld2)                       # deterministic
        ys, lds = seq(x[4096:4096 + 777])
        assert torch.equal(ys, y[4096:4096 + 777]) and torch.equal(lds, ld[4096:4096 + 777])   # tile-position invariant
    assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(ld).all())
    # Four stacked layers: a sample whose intermediate y lands within the bf16 error of the +-pi seam is
    # wrapped to the other side and then follows a different (equally valid) branch of the NEXT layer's
    # conditioner, so agreement is stated for the bulk of the samples.
    d = _circ(y, y32).max(dim=1).values
    assert float(d.median()) < 5e-3
    assert float((d < 5e-2).float().mean()) > 0.97
    assert float((ld - ld32).abs().median()) < 2e-2


@pytest.mark.parametrize('batch', [1, 300, 128 * 148 + 77])
def test_chain_launch_equals_layer_by_layer(batch):
    """SequentialFlow runs the four bf16 layers as ONE launch (work items layer-major, per-tile flags):
    bit-identical to one launch per layer, including the ragged last tile."""
    seq, _ = cfg_flow_modules('cfg2', DEV)
    x = cases.cfg_input('cfg2', batch).to(DEV)
    for maf in seq:
        maf.precision = 'bf16'
    with torch.no_grad():
        y, ld = seq(x)
        y2, ld2 = seq(x)
        cur, tot = x, None
        for maf in seq:
            cur, l = maf(cur)
            tot = l if tot is None else tot + l
    assert torch.equal(y, y2) and torch.equal(ld, ld2)
    assert torch.equal(y, cur)
    assert float((ld - tot).abs().max()) < 1e-5          # same per-layer values, summed in the same order
    assert int(seq[0]._fused._tables(torch.device(DEV))['err'].item()) == 0


@pytest.mark.parametrize('batch', [1, 129, 1000])
def test_fused_inverse_layers(batch):
    """tfepb_maf_spline_inverse_bf16, one layer (ascending and descending degrees): against the EXACT inverse
    sweep of the same y (differences = bf16 operand rounding: same stated tolerance as the forward kernel)
    and as the inverse of the fused forward (same roundings in both directions: tight, except for the few
    samples whose x sits on a bf16 rounding boundary and flips the conditioner input)."""
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=2)
    x = cases.cfg_input('cfg2', batch).to(DEV)
    for maf in seq:
        with torch.no_grad():
            maf.precision = 'bf16'
            y, ld = maf(x)
            xi, ldi = maf.inverse(y)
            xi2, ldi2 = maf.inverse(y)
            maf.precision = 'fp32'
            xe, lde = maf.inverse(y)
        assert torch.equal(xi, xi2) and torch.equal(ldi, ldi2)
        assert xi.shape == (batch, 66) and ldi.shape == (batch,)
        assert float(_circ(xi, xe).max()) < 5e-2 and float(_circ(xi, xe).mean()) < 2e-3
        assert float((ldi - lde).abs().max()) < 1e-1 and float((ldi - lde).abs().mean()) < 8e-3
        d = _circ(xi, x).max(dim=1).values
        assert float(d.median()) < 2e-5
        assert float((d < 2e-2).float().mean()) > 0.97
        assert float((ld + ldi).abs().median()) < 5e-5


def test_fused_inverse_chain_round_trip():
    """SequentialFlow.inverse with four bf16 layers = ONE launch of the inverse kernel (layers in reverse
    order, per-tile flags), bit-identical to one launch per layer; round trip through the fused forward."""
    seq, _ = cfg_flow_modules('cfg2', DEV)
    x = cases.cfg_input('cfg2', 128 * 148 + 77).to(DEV)
    for maf in seq:
        maf.precision = 'bf16'
    with torch.no_grad():
        y, ld = seq(x)
        xi, ldi = seq.inverse(y)
        cur, tot = y, None
        for maf in reversed(seq):
            cur, l = maf.inverse(cur)
            tot = l if tot is None else tot + l
    assert torch.equal(xi, cur)
    assert float((ldi - tot).abs().max()) < 1e-5
    d = _circ(xi, x).max(dim=1).values
    assert float(d.median()) < 1e-4
    assert float((d < 5e-2).float().mean()) > 0.9
    assert float((ld + ldi).abs().median()) < 2e-4
    assert bool(torch.isfinite(xi).all()) and bool(torch.isfinite(ldi).all())
    assert int(seq[0]._fused._tables(torch.device(DEV))['err'].item()) == 0


def test_long_chain_is_split_into_launches():
    """Six layers: forward runs as 4 + 2 layers per launch (shared-memory budget of the feature tables), inverse
    as one launch; both equal the layer-by-layer results."""
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=6)
    x = cases.cfg_input('cfg2', 700).to(DEV)
    for maf in seq:
        maf.precision = 'bf16'
    with torch.no_grad():
        y, ld = seq(x)
        cur, tot = x, None
        for maf in seq:
            cur, l = maf(cur)
            tot = l if tot is None else tot + l
        xi, ldi = seq.inverse(y)
    assert torch.equal(y, cur) and float((ld - tot).abs().max()) < 1e-5
    d = _circ(xi, x).max(dim=1).values
    assert float(d.median()) < 1e-4 and float((ld + ldi).abs().median()) < 5e-4


@pytest.mark.parametrize('cfg,D', [('cfg2mix', 66), ('cfg5', 66), ('cfg2mixemb', 48)])
def test_mixed_and_non_circular_splines(cfg, D):
    """The generic spline epilogue of the fused kernels: a MixedTransformer of circular (torsions) and ordinary
    splines with linear tails (Cartesians) -- BASELINE.json's "dihedral / Cartesian mix" -- a purely
    non-circular flow (cfg5's transformer at D = 66), and the mix with a PeriodicEmbedding in front of the
    conditioner (what the reference's MixedMAFMap runs, app/mixedmaf.py:341-353: the (cos, sin) lift is fused into
    the operand staging of both kernels).  Forward against the exact fp32 path (stated bf16 tolerance),
    inverse as a round trip, chain launch equal to layer-by-layer launches."""
    seq, flows = cfg_flow_modules(cfg, DEV, n_layers=3, D=D)
    x = cases.cfg_input(cfg, 1500, D=D).to(DEV)
    if cfg == 'cfg5':
        x = x * 2.5                                   # a good share of the samples in the tails beyond +-5
    period = torch.full((D,), float('inf'))
    if cfg != 'cfg5':
        period[[f for f in range(D) if f % 3 == 2]] = 2 * math.pi
    if cfg == 'cfg2mixemb':
        assert all(m._embedding is not None for m in seq)

    def dist(a, b):
        d = (a.double().cpu() - b.double().cpu()).abs()
        return torch.minimum(d, (period - d).abs())

    with torch.no_grad():
        y32, ld32 = seq[0](x)
        for maf in seq:
            maf.precision = 'bf16'
        y, ld = seq[0](x)
        assert float(dist(y, y32).max()) < 5e-2 and float(dist(y, y32).mean()) < 2e-3
        assert float((ld - ld32).abs().max()) < 1e-1 and float((ld - ld32).abs().mean()) < 8e-3
        xi, ldi = seq[0].inverse(y)
        d = dist(xi, x).max(dim=1).values
        assert float(d.median()) < 2e-5 and float((d < 2e-2).float().mean()) > 0.97
        assert float((ld + ldi).abs().median()) < 5e-5
        # chain: one launch == layer by layer, in both directions
        yc, ldc = seq(x)
        cur, tot = x, None
        for maf in seq:
            cur, l = maf(cur)
            tot = l if tot is None else tot + l
        assert torch.equal(yc, cur) and float((ldc - tot).abs().max()) < 1e-5
        xc, ldci = seq.inverse(yc)
        dc = dist(xc, x).max(dim=1).values
        assert float(dc.median()) < 1e-4 and float((ldc + ldci).abs().median()) < 2e-4
    if cfg == 'cfg5':
        assert float((x.abs() > 5).float().mean()) > 0.01          # the tails were exercised


@pytest.mark.parametrize('cfg,D', [('cfg2mixemb', 46), ('cfg2mixemb', 33), ('cfg2mix', 12), ('cfg2mix', 3)])
def test_narrow_and_odd_widths(cfg, D):
    """Odd numbers of conditioner inputs (the constant-one columns then straddle two bf16 pair columns), lifted
    (cos, sin) pairs next to plain features, and layers too narrow for the two-half hand-over of the hidden
    layers (one half, one output chunk): forward against the exact path, inverse as a round trip."""
    seq, _ = cfg_flow_modules(cfg, DEV, n_layers=2, D=D)
    x = cases.cfg_input(cfg, 700, D=D).to(DEV)
    period = torch.full((D,), float('inf'))
    period[[f for f in range(D) if f % 3 == 2]] = 2 * math.pi

    def dist(a, b):
        d = (a.double().cpu() - b.double().cpu()).abs()
        return torch.minimum(d, (period - d).abs())

    with torch.no_grad():
        for maf in seq:
            y32, ld32 = maf(x)
            maf.precision = 'bf16'
            assert maf._fused_plan() is not None, maf._fused_why
            y, ld = maf(x)
            xi, ldi = maf.inverse(y)
            assert float(dist(y, y32).max()) < 5e-2 and float(dist(y, y32).mean()) < 2e-3
            assert float((ld - ld32).abs().mean()) < 8e-3
            d = dist(xi, x).max(dim=1).values
            assert float(d.median()) < 2e-5 and float((ld + ldi).abs().median()) < 5e-5


def test_mixed_maf_map_shape_with_explicit_hidden_widths():
    """The shape the reference's MixedMAFMap builds at D = 66 (22 torsions lifted to (cos, sin), circular splines on them,
    ordinary splines on the 44 Cartesians): with the default hidden width (381) the layers fall outside the
    tensor-memory plan, with ``hidden_layers=[334, 334]`` the whole chain runs on the fused kernels in both directions."""
    seq, _ = cfg_flow_modules('cfg2mixemb', DEV, n_layers=4, D=66, hidden_layers=[334, 334])
    x = cases.cfg_input('cfg2mixemb', 4000, D=66).to(DEV)
    period = torch.full((66,), float('inf'))
    period[[f for f in range(66) if f % 3 == 2]] = 2 * math.pi

    def dist(a, b):
        d = (a.double().cpu() - b.double().cpu()).abs()
        return torch.minimum(d, (period - d).abs())

    with torch.no_grad():
        y32, ld32 = seq[0](x)
        for maf in seq:
            maf.precision = 'bf16'
            assert maf._fused_plan() is not None and maf._fused_plan().inverse_eligibility(maf) is None, maf._fused_why
        y, ld = seq[0](x)
        assert float(dist(y, y32).max()) < 5e-2 and float(dist(y, y32).mean()) < 2e-3
        assert float((ld - ld32).abs().mean()) < 8e-3
        yc, ldc = seq(x)                              # one launch for the four layers
        xc, ldci = seq.inverse(yc)
        d = dist(xc, x).max(dim=1).values
        assert float(d.median()) < 1e-4 and float((ldc + ldci).abs().median()) < 2e-4
        assert int(seq[0]._fused._tables(torch.device(DEV))['err'].item()) == 0


@pytest.mark.parametrize('cfg,D', [('cfg2cond', 66), ('cfg2condemb', 36)])
def test_conditioning_features(cfg, D):
    """Features of degree -1 (conditioning atoms) feed the conditioner and pass through.  Forward: no slot, no output
    rows.  Inverse: they are staged into the x operand before the sweep and leading steps without a feature bring
    in the hidden units that depend on them alone (30 units of each layer at D = 66: chunks of 15).  With a
    PeriodicEmbedding the conditioning features are lifted to (cos, sin) as well."""
    seq, _ = cfg_flow_modules(cfg, DEV, n_layers=3, D=D)
    x = cases.cfg_input(cfg, 1000, D=D).to(DEV)
    with torch.no_grad():
        y32, ld32 = seq[0](x)
        s32, sl32 = seq(x)
        for maf in seq:
            maf.precision = 'bf16'
            assert maf._fused_plan() is not None, maf._fused_why
            assert maf._fused_plan().inverse_eligibility(maf) is None
        y, ld = seq[0](x)
        assert torch.equal(y[:, :6], x[:, :6])
        assert float(_circ(y, y32).max()) < 5e-2 and float(_circ(y, y32).mean()) < 2e-3
        assert float((ld - ld32).abs().mean()) < 8e-3
        xi, ldi = seq[0].inverse(y)
        d = _circ(xi, x).max(dim=1).values
        assert torch.equal(xi[:, :6], x[:, :6])
        assert float(d.median()) < 2e-5 and float((d < 2e-2).float().mean()) > 0.97
        assert float((ld + ldi).abs().median()) < 5e-5
        seq[0].precision = 'fp32'
        xe, lde = seq[0].inverse(y)                  # exact sweep of the same y
        seq[0].precision = 'bf16'
        assert float(_circ(xi, xe).mean()) < 2e-3 and float((ldi - lde).abs().mean()) < 8e-3
        from tfep_b200 import _fused
        plans = [maf._fused_plan() for maf in seq]
        assert _fused.chain_compatible(plans)        # one launch even where the layers split their hidden layers differently
        yc, ldc = seq(x)
        cur, tot = x, None
        for maf in seq:
            cur, l = maf(cur)
            tot = l if tot is None else tot + l
        assert torch.equal(yc, cur) and float((ldc - tot).abs().max()) < 1e-5
        assert float(_circ(yc, s32).max(dim=1).values.median()) < 5e-3
        xc, ldci = seq.inverse(yc)
        dc = _circ(xc, x).max(dim=1).values
        assert float(dc.median()) < 1e-4 and float((ldc + ldci).abs().median()) < 2e-4
        assert int(seq[0]._fused._tables(torch.device(DEV))['err'].item()) == 0


def test_layers_beyond_the_tensor_memory_plan_use_the_general_gemm():
    """D = 66 with 22 lifted torsions asks for hidden layers of 381 units: more than the tensor-memory plan of the
    one-launch kernel holds.  precision='bf16' then runs the conditioner on the general tensor-core GEMM (and the
    inverse on the exact sweep) instead of failing."""
    seq, _ = cfg_flow_modules('cfg2mixemb', DEV, n_layers=1, D=66)
    maf = seq[0]
    x = cases.cfg_input('cfg2mixemb', 300, D=66).to(DEV)
    with torch.no_grad():
        y32, ld32 = maf(x)
        maf.precision = 'bf16'
        assert maf._fused_plan() is None and 'tensor-memory plan' in maf._fused_why
        y, ld = maf(x)
        xi, ldi = maf.inverse(y)
    assert float((ld - ld32).abs().mean()) < 8e-3
    # forward with bf16 operands, inverse exact: the round trip closes to the stated bf16 tolerance
    d = (xi - x).abs()
    d = torch.minimum(d, (2 * math.pi - d).abs())
    assert float(d.max()) < 5e-2 and float(d.mean()) < 2e-3


def test_bf16_outside_the_fused_kernel_uses_the_tensor_core_gemm():
    """precision='bf16' where the one-launch kernel does not apply -- training (autograd) and other transformers --
    runs the MADE conditioner on the general tensor-core GEMM, forward and backward, with the exact transformer
    kernels: results and gradients agree with the fp32 path to bf16 operand rounding."""
    from tfep_b200.nn.conditioners import generate_degrees
    from tfep_b200.nn.flows import MAF
    # (a) spline flow under autograd: gradients w.r.t. x and every parameter
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=1)
    maf = seq[0]
    x = cases.cfg_input('cfg2', 300).to(DEV)
    cy, cl = cases.normal((300, 66), 78).to(DEV), cases.normal((300,), 79).to(DEV)

    def grads(precision):
        maf.precision = precision
        maf.zero_grad(set_to_none=True)
        xg = x.clone().requires_grad_(True)
        y, ld = maf(xg)
        ((y * cy).sum() + (ld * cl).sum()).backward()
        return y.detach(), ld.detach(), xg.grad, {k: p.grad.clone() for k, p in maf.named_parameters()}

    y32, ld32, gx32, gp32 = grads('fp32')
    y16, ld16, gx16, gp16 = grads('bf16')
    assert float(_circ(y16, y32).mean()) < 2e-3 and float((ld16 - ld32).abs().mean()) < 8e-3

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-12))

    assert rel(gx16, gx32) < 5e-2
    for k in gp32:
        assert rel(gp16[k], gp32[k]) < 5e-2, k
    # (b) a transformer the fused kernel does not cover (affine), inference
    affine = MAF(generate_degrees(6), initialize_identity=False).to(DEV)
    xa = torch.randn(40, 6, device=DEV)
    with torch.no_grad():
        ya, lda = affine(xa)
        affine.precision = 'bf16'
        yb, ldb = affine(xa)
    assert float((ya - yb).abs().max()) < 5e-2 and float((lda - ldb).abs().max()) < 5e-2


def test_empty_batches_through_the_fused_paths():
    """B = 0 is legal everywhere (the reference's flows accept it): chain forward / inverse, one layer, and the wrapper
    flows' fused pre / post kernels."""
    from tfep_b200.nn.flows import CenteredCentroidFlow, OrientedFlow
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=2)
    for maf in seq:
        maf.precision = 'bf16'
    x = torch.empty(0, 66, device=DEV)
    with torch.no_grad():
        for fn in (seq, seq.inverse, seq[0], seq[0].inverse):
            y, ld = fn(x)
            assert y.shape == (0, 66) and ld.shape == (0,)
        wrapped = CenteredCentroidFlow(OrientedFlow(seq), space_dimension=3).to(DEV)
        y, ld = wrapped(torch.empty(0, 72, device=DEV))
        assert y.shape == (0, 72) and ld.shape == (0,)


def test_protocol_stress_is_deterministic():
    """compute-sanitizer is closed on this pool (profiles/README.md), so the hand-rolled synchronisation of the chain
    kernels -- mbarrier rings, TMEM hand-overs, inter-CTA tile flags with per-launch epochs, cooperative co-residency -- is
    exercised the hard way: 150 back-to-back launches over batch sizes that give ragged tiles, fewer tiles than SMs, one
    tile per SM and many rounds, forward and inverse, every result bit-identical to the first run of its shape (a lost or
    early hand-over shows up as a wrong or non-reproducible tile; a dead-lock trips the kernels' clock watchdog)."""
    seq, _ = cfg_flow_modules('cfg2', DEV)
    for maf in seq:
        maf.precision = 'bf16'
    shapes = [1, 127, 129, 128 * 37 + 5, 128 * 148, 128 * 148 + 1, 128 * 400 + 77, 65536]
    xs = {b: cases.cfg_input('cfg2', b).to(DEV) for b in shapes}
    first = {}
    with torch.no_grad():
        for it in range(150):
            b = shapes[(it * 5) % len(shapes)]
            y, ld = seq(xs[b])
            if it % 3 == 0:
                xi, ldi = seq.inverse(y)
            else:
                xi = ldi = None
            if b not in first:
                first[b] = [y.clone(), ld.clone(), None, None]
            ref = first[b]
            assert torch.equal(y, ref[0]) and torch.equal(ld, ref[1]), (it, b)
            if xi is not None:
                if ref[2] is None:
                    ref[2], ref[3] = xi.clone(), ldi.clone()
                assert torch.equal(xi, ref[2]) and torch.equal(ldi, ref[3]), (it, b)
    torch.cuda.synchronize()
