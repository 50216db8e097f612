"""Host-side logic of the drop-in modules: construction, degrees, masks, state_dict layout, packing.

No kernels run here (no GPU): these tests cover what the reference's own constructor-level tests cover
(tfep/tests/nn/conditioners/test_made.py:31-119, tests/nn/flows/test_maf.py) plus the packing plan.
"""

import json
import os

import pytest
import torch

from helpers import GOLDEN, to_maf, to_module
from oracle import cases
from oracle import flow_oracle as fo
from tfep_b200._lib import TfepB200Error
from tfep_b200._pack import MadePlan
from tfep_b200.nn import masked
from tfep_b200.nn.conditioners import MADE, generate_degrees
from tfep_b200.nn.flows import MAF, SequentialFlow
from tfep_b200.nn.transformers import (AffineTransformer, MixedTransformer, MoebiusTransformer,
                                       NeuralSplineTransformer, SOSPolynomialTransformer)

TAB = json.load(open(os.path.join(GOLDEN, 'degrees.json')))


def test_generate_degrees_known_answers():
    for row in TAB['generate_degrees']:
        assert generate_degrees(row['n_features'], **row['kwargs']).tolist() == row['expected']
    with pytest.raises(ValueError, match='Accepted string values for'):
        generate_degrees(n_features=2, order='wrong')


def test_hidden_degrees_masks_and_parameter_count():
    for row in TAB['hidden_degrees']:
        din, dout = torch.tensor(row['degrees_in']), torch.tensor(row['degrees_out'])
        got = MADE._get_degrees_hidden(din, dout, row['hidden_layers'])
        assert [h.tolist() for h in got] == row['expected']
        made = MADE(din, dout, row['hidden_layers'])
        assert [int(l.mask.sum()) for l in made.layers[::2]] == row['mask_sums']
        assert int(made.n_parameters()) == row['n_parameters']
        assert made.dimension_in == len(din) and made.dimension_out == len(dout)
        assert made.dimensions_hidden.tolist() == [len(h) for h in row['expected']]


def test_made_errors():
    with pytest.raises(ValueError, match='is too small for the number of input features'):
        MADE(torch.tensor([0, 1, 2, 3]), torch.tensor([0, 1, 2, 3]), hidden_layers=[2])
    with pytest.raises(ValueError, match='nodes with degrees that will be ignored'):
        MADE(torch.tensor([0, 1, 2]), torch.tensor([0, 1, 2]), hidden_layers=[[0, 2, 1]])


def test_config_shapes_match_reference():
    """Layer widths and mask non-zeros of the BASELINE.json configurations."""
    mk = {
        'cfg1': (66, AffineTransformer(), 1),
        'cfg2': (66, NeuralSplineTransformer(torch.zeros(66), torch.ones(66), 8, circular=True), 1),
        'cfg3_sos': (300, SOSPolynomialTransformer(2), 1),
        'cfg3_moebius': (300, MoebiusTransformer(3), 3),
    }
    for name, (D, t, rep) in mk.items():
        maf = MAF(generate_degrees(D, repeats=rep), transformer=t)
        lin = list(maf._conditioner.layers[::2])
        assert [lin[0].in_features] + [l.out_features for l in lin] == TAB['config_shapes'][name]['dims']
        assert [int(l.mask.sum()) for l in lin] == TAB['config_shapes'][name]['nnz']
        # the packed plan keeps exactly the reference's non-zeros
        assert maf._pack()['plan'].nnz == TAB['config_shapes'][name]['nnz']


def test_state_dict_keys_are_the_reference_ones():
    maf = MAF(generate_degrees(6), NeuralSplineTransformer(torch.zeros(6), torch.ones(6), 4, circular=True))
    keys = set(maf.state_dict().keys())
    expected = {'_transformer_indices', '_inverse_masks', '_fixed_indices', '_conditioner_indices'}
    expected |= {f'_conditioner.layers.{i}.{k}' for i in (0, 2, 4) for k in ('bias', 'weight_g', 'weight_v', 'mask')}
    expected |= {f'_transformer.{k}' for k in ('x0', 'xf', 'n_bins', '_y0', '_yf', '_circular',
                                               '_identity_boundary_slopes', '_learn_lower_bound',
                                               '_learn_upper_bound', '_min_bin_size', '_min_slope')}
    assert keys == expected
    nown = MAF(generate_degrees(4), weight_norm=False)
    assert '_conditioner.layers.0.weight' in nown.state_dict()


def test_maf_degree_validation_and_buffers():
    with pytest.raises(ValueError, match='degrees_in must assume consecutive values'):
        MAF(degrees_in=[0, 2, 3])
    with pytest.raises(ValueError, match='degrees_in must assume consecutive values'):
        MAF(degrees_in=[1, 2])
    maf = MAF(degrees_in=[-1, 1, 0, -1, 1])
    assert maf._fixed_indices.tolist() == [0, 3] and maf._transformer_indices.tolist() == [1, 2, 4]
    assert maf._inverse_masks.tolist() == [[False, False, True, False, False], [False, True, False, False, True]]
    assert maf.has_fixed_indices
    full = MAF(degrees_in=[0, 1, 2])
    assert len(full._transformer_indices) == 0 and not full.has_fixed_indices


def test_identity_initialisation_parameters():
    """initialize_identity zeroes the last layer's g and loads the identity parameters in its bias
    (reference autoregressive.py:133-137, made.py:358-364)."""
    for t in (AffineTransformer(), SOSPolynomialTransformer(3), MoebiusTransformer(2),
              NeuralSplineTransformer(torch.full((4,), -1.), torch.full((4,), 1.), 5)):
        maf = MAF(generate_degrees(4, repeats=2 if isinstance(t, MoebiusTransformer) else 1), transformer=t)
        last = maf._conditioner.layers[-1]
        assert float(last.weight_g.abs().max()) == 0.0
        assert torch.equal(last.bias.data, t.get_identity_parameters(4).to(last.bias))


def test_transformer_host_api_matches_oracle():
    for name, (spec, n, x, par) in cases.transformer_cases(torch.float32).items():
        mod = to_module(spec)
        torch.manual_seed(3)          # (the symmetrized Moebius identity is a tiny random tensor)
        ident = mod.get_identity_parameters(n)
        torch.manual_seed(3)
        assert torch.equal(ident, spec.identity_params(n)), name
        deg = fo.gen_degrees(n, repeats=spec.dimension) if isinstance(spec, (fo.Moebius, fo.SymMoebius)) else fo.gen_degrees(n)
        assert torch.equal(mod.get_degrees_out(deg), spec.degrees_out(deg)), name
        parts = mod._parts(n)
        cols = torch.cat([p.ref_columns().flatten() for p in parts]).sort().values
        assert torch.equal(cols, torch.arange(len(spec.identity_params(n)))), name


def test_spline_constructor_errors():
    x0, xf = torch.zeros(2), torch.ones(2)
    with pytest.raises(ValueError, match='circular spline with learnable limits'):
        NeuralSplineTransformer(x0, xf, 3, circular=True, learn_lower_bound=True)
    with pytest.raises(ValueError, match='minimum bin size'):
        NeuralSplineTransformer(x0, xf, 3, min_bin_size=0.0)
    with pytest.raises(ValueError, match='minimum slope'):
        NeuralSplineTransformer(x0, xf, 3, min_slope=1.0)
    assert NeuralSplineTransformer(x0, xf, 8).n_parameters_per_feature == 25
    assert NeuralSplineTransformer(x0, xf, 8, circular=True).n_parameters_per_feature == 25
    assert NeuralSplineTransformer(x0, xf, 8, identity_boundary_slopes=True).n_parameters_per_feature == 23
    assert NeuralSplineTransformer(x0, xf, 8, circular=True, identity_boundary_slopes=True).n_parameters_per_feature == 24
    assert NeuralSplineTransformer(x0, xf, 8, learn_lower_bound=True, learn_upper_bound=True).n_parameters_per_feature == 27
    with pytest.raises(ValueError, match='strictly greater than 1'):
        SOSPolynomialTransformer(1)
    with pytest.raises(ValueError, match='greater than 1'):
        MixedTransformer([AffineTransformer()], [[0]])


def test_effective_weight_matches_oracle_and_masks_gradients():
    torch.manual_seed(0)
    mask = (torch.rand(7, 5) > 0.4).float()
    mask[2] = 0.0                                     # a fully masked row: 0/0 in the naive formula
    v = (torch.randn(7, 5) * mask).requires_grad_(True)
    g = torch.rand(7, 1).requires_grad_(True)
    w = masked.effective_weight(v, g, mask)
    assert torch.allclose(w, fo.effective_weight(v.detach(), g.detach(), mask), atol=1e-7)
    (w * torch.randn(7, 5)).sum().backward()
    assert torch.isfinite(v.grad).all() and torch.isfinite(g.grad).all()
    assert float(v.grad[mask == 0].abs().max()) == 0.0 and float(g.grad[2].abs()) == 0.0


def test_packing_plan_is_a_relabelling():
    """Packed weights are a permutation of the reference ones and the skipped ranges hold only zeros."""
    for name, case in cases.maf_cases(torch.float32).items():
        oracle, sd = cases.build_oracle(case)
        maf = to_maf(case, sd)
        pk = maf._pack()
        plan = pk['plan']
        ws, bs = zip(*maf._conditioner.effective_weights())
        pw, pb = plan.pack([w.detach() for w in ws], [b.detach() for b in bs])
        for l, (w_ref, b_ref) in enumerate(oracle.layers):
            assert torch.allclose(ws[l].detach(), w_ref, atol=1e-6), name
            back = torch.empty_like(w_ref)
            rows, cols = plan.perms[l + 1], plan.perms[l]
            back[rows[:, None], cols[None, :]] = pw[l]
            assert torch.allclose(back, w_ref, atol=1e-6), name
            for t, (kb, ke) in enumerate(plan.k_ranges[l].tolist()):
                tile = pw[l][t * 64:(t + 1) * 64]
                assert float(tile[:, :kb].abs().sum()) == 0.0 and float(tile[:, ke:].abs().sum()) == 0.0, name
            for t, (nb, ne) in enumerate(plan.n_ranges[l].tolist()):
                tile = pw[l][:, t * 64:(t + 1) * 64]
                assert float(tile[:nb].abs().sum()) == 0.0 and float(tile[ne:].abs().sum()) == 0.0, name
        # every degree group owns a contiguous block of packed output rows; together they tile the output
        pos = 0
        for grp in pk['groups']:
            assert grp['rows'][0] == pos
            pos = grp['rows'][1]
        assert pos == maf._conditioner.dimension_out


def test_padded_chunk_layout_of_the_fused_transformer_epilogue():
    """tfep_b200/_txfused.py: the output layer of an SOS / Moebius / affine MAF in 16-column chunks of whole units (host logic):
    the padded plan is the packed plan with all-zero rows in the pad columns, units stay in degree order (the staircase
    ranges of the padded layer never exceed those of the unpadded rows they hold) and the column table lists the x columns
    of every unit."""
    from tfep_b200 import _txfused
    from tfep_b200.nn.transformers import MoebiusTransformer, SOSPolynomialTransformer
    torch.manual_seed(0)
    for maf, upc, ppu in (
            (MAF(generate_degrees(40, order='descending'), SOSPolynomialTransformer(2), initialize_identity=False), 3, 5),
            (MAF(generate_degrees(33, repeats=3), MoebiusTransformer(dimension=3), initialize_identity=False), 5, 3),
            (MAF(generate_degrees(19), initialize_identity=False), 8, 2)):
        pk = maf._pack()
        assert _txfused.eligibility(maf, pk) is None
        txp = _txfused.TcTxPlan(maf, pk)
        made = maf._conditioner
        with torch.no_grad():
            pw, pb = made.packed_weights(pk['plan'])
            qw, qb = made.packed_weights(txp.plan)
        n_rows = pw[-1].shape[0]
        n_units = n_rows // ppu
        assert txp.n_padded == -(-n_units // upc) * 16 == qw[-1].shape[0]
        for u in range(n_units):
            r0 = (u // upc) * 16 + (u % upc) * ppu
            assert torch.equal(qw[-1][r0:r0 + ppu], pw[-1][u * ppu:(u + 1) * ppu])
            assert torch.equal(qb[-1][r0:r0 + ppu], pb[-1][u * ppu:(u + 1) * ppu])
        pad = torch.ones(txp.n_padded, dtype=torch.bool)
        for u in range(n_units):
            r0 = (u // upc) * 16 + (u % upc) * ppu
            pad[r0:r0 + ppu] = False
        assert float(qw[-1][pad].abs().sum()) == 0.0 and float(qb[-1][pad].abs().sum()) == 0.0
        for l in range(len(pw) - 1):
            assert torch.equal(qw[l], pw[l])
        # x columns of the units in chunk order = the columns of the degree-sorted features
        part = pk['parts'][0]
        order = torch.argsort(pk['bases'][0].view(-1, 3 if ppu == 3 else 1)[:, 0])
        cols = part.x_columns().view(-1, 3 if ppu == 3 else 1)[order].reshape(-1)
        assert torch.equal(txp.cols.long(), cols)
        # staircase: everything outside the forward k-block range of a 256-row tile of the padded layer is zero
        fwd = txp.plan.tc_ranges('cpu')[0][-1]
        for t, (kb, ke) in enumerate(fwd.tolist()):
            tile = qw[-1][t * 256:(t + 1) * 256]
            assert float(tile[:, :kb * 64].abs().sum()) == 0.0 and float(tile[:, ke * 64:].abs().sum()) == 0.0
    # flows the epilogue does not cover say why
    sos3 = MAF(generate_degrees(8), SOSPolynomialTransformer(3), initialize_identity=False)
    assert 'polynomials' in _txfused.eligibility(sos3, sos3._pack())
    cond = MAF([-1] + generate_degrees(6).tolist(), initialize_identity=False)      # conditioning features pass through
    assert _txfused.eligibility(cond, cond._pack()) is None and _txfused.TcTxPlan(cond, cond._pack()).passthrough


def test_staircase_skips_about_half_of_the_headline_config():
    maf = MAF(generate_degrees(66), NeuralSplineTransformer(torch.zeros(66), torch.ones(66), 8, circular=True))
    plan = maf._pack()['plan']
    dense = [66 * 328, 328 * 328, 328 * 1650]
    scheduled = [sum((ke - kb) * min(64, n - 64 * t) for t, (kb, ke) in enumerate(r.tolist()))
                 for r, n in zip(plan.k_ranges, (328, 328, 1650))]
    assert sum(scheduled) < 0.62 * sum(dense)
    assert plan.masked_macs == 338277


def test_no_cpu_fallback():
    """The product path refuses CPU tensors loudly instead of silently computing elsewhere."""
    maf = MAF(generate_degrees(4), initialize_identity=False)
    with pytest.raises(TfepB200Error, match='no CPU fallback'):
        maf(torch.randn(3, 4))
    from tfep_b200.analysis import fep_estimator
    with pytest.raises(TfepB200Error, match='no CPU fallback'):
        fep_estimator(torch.randn(10))


def test_sequential_flow_container():
    seq = SequentialFlow(MAF(generate_degrees(3)), MAF(generate_degrees(3, order='descending')))
    assert int(seq.n_parameters()) == sum(int(f.n_parameters()) for f in seq)


@pytest.mark.parametrize('prec', ['f32', 'f64'])
def test_wrapper_flows_on_host(prec):
    """PartialFlow / CenteredCentroidFlow / OrientedFlow (device-agnostic tensor algebra either side of the MAF
    kernels) around an oracle-backed CPU flow: outputs, inverse and the gradient w.r.t. x against the reference's
    golden vectors, plus the constraints the wrappers exist for."""
    from helpers import OracleFlowModule, golden, rel_err, to_wrapper
    dtype = {'f32': torch.float32, 'f64': torch.float64}[prec]
    tol = {'f32': 1e-5, 'f64': 1e-11}[prec]
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        g = golden(f'wrappers_{prec}.npz')
        for name, case in cases.wrapper_cases(dtype).items():
            inner, _ = cases.build_oracle(case['inner'], dtype)
            flow = to_wrapper(case, OracleFlowModule(inner), dtype)
            x = case['x'].clone().requires_grad_(True)
            y, ld = flow(x)
            assert rel_err(y, g[f'{name}/y']) < tol and rel_err(ld, g[f'{name}/ld']) < tol, name
            cy, cl = cases.normal(tuple(y.shape), 78, dtype), cases.normal(tuple(ld.shape), 79, dtype)
            ((y * cy).sum() + (ld * cl).sum()).backward()
            assert rel_err(x.grad, g[f'{name}/gx']) < 20 * tol, name
            if case['invertible']:
                with torch.no_grad():
                    xi, ldi = flow.inverse(torch.from_numpy(g[f'{name}/y']))
                assert rel_err(xi, g[f'{name}/xinv']) < 20 * tol and rel_err(ldi, g[f'{name}/ldinv']) < 20 * tol, name
                assert rel_err(xi, case['x']) < 1e3 * tol and rel_err(ld.detach() + ldi, torch.zeros_like(ldi)) < 1e3 * tol, name
        # the constraints: fixed features untouched; centroid preserved; constrained coordinates stay zero in the frame
        case = cases.wrapper_cases(dtype)['partial']
        y = torch.from_numpy(g['partial/y'])
        assert torch.equal(y[:, [1, 4]], case['x'][:, [1, 4]])
        case = cases.wrapper_cases(dtype)['centroid']
        yc = torch.from_numpy(g['centroid/y']).reshape(12, 4, 3).mean(dim=1)
        assert rel_err(yc, case['x'].reshape(12, 4, 3).mean(dim=1)) < 10 * tol
    finally:
        torch.set_default_dtype(old)


def test_wrapper_flow_constructor_errors():
    from tfep_b200.nn.flows import CenteredCentroidFlow, OrientedFlow
    inner = torch.nn.Identity()
    with pytest.raises(ValueError, match="'return_partial=True' is supported only if 'translate_back=False'"):
        CenteredCentroidFlow(inner, 3, return_partial=True)
    with pytest.raises(ValueError, match="'origin' must have length equal to 'space_dimension'"):
        CenteredCentroidFlow(inner, 3, origin=[0.0, 1.0])
    with pytest.raises(ValueError, match="'weights' must have the same length as 'subset_point_indices'"):
        CenteredCentroidFlow(inner, 3, subset_point_indices=[0, 1], weights=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError, match="'return_partial=True' is supported only if 'rotate_back=False'"):
        OrientedFlow(inner, return_partial=True)
    with pytest.raises(ValueError, match="must be different"):
        OrientedFlow(inner, axis_point_idx=1, plane_point_idx=1)
    with pytest.raises(ValueError, match="must be constrained on an axis on the same plane"):
        OrientedFlow(inner, axis='z', plane='xy')
    with pytest.raises(ValueError, match="only if 'translate_back' is set to True"):
        CenteredCentroidFlow(inner, 3, translate_back=False).inverse(torch.zeros(2, 6))
    with pytest.raises(ValueError, match="only if 'rotate_back' is set to True"):
        OrientedFlow(inner, rotate_back=False).inverse(torch.zeros(2, 9))


def _load_embedding(name, device='cpu'):
    """tfep_b200 embedding of an oracle.cases.embedding_cases entry with the reference's seeded parameters."""
    import types
    import tfep_b200.nn.embeddings as E
    from helpers import golden
    g = golden('embeddings.npz')
    build, n, deg = cases.embedding_cases()[name]
    emb = build(types.SimpleNamespace(FlipInvariantEmbedding=E.FlipInvariantEmbedding, MixedEmbedding=E.MixedEmbedding,
                                      PeriodicEmbedding=E.PeriodicEmbedding))
    sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f'{name}/sd/')}
    missing, unexpected = emb.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if 'weight' in k or 'bias' in k], (missing, unexpected)
    return emb.to(device), g, deg


@pytest.mark.parametrize('name', ['flip', 'mixed_flips'])
def test_learnable_embeddings_match_reference_golden(name):
    """FlipInvariantEmbedding / MixedEmbedding (reference mafembed.py:174-446): same parameter names, outputs, output
    degrees and input gradients as the reference; invariant to flipping the sign of an embedded vector."""
    from helpers import rel_err
    emb, g, deg = _load_embedding(name)
    x = torch.from_numpy(g[f'{name}/x']).requires_grad_(True)
    y = emb(x)
    assert rel_err(y, g[f'{name}/y']) < 1e-6
    (y * cases.normal(tuple(y.shape), 42)).sum().backward()
    assert rel_err(x.grad, g[f'{name}/gx']) < 1e-5
    assert torch.equal(emb.get_degrees_out(deg), torch.from_numpy(g[f'{name}/deg']))
    if name == 'flip':
        xf = x.detach().clone()
        xf[:, [1, 2, 3, 4]] *= -1
        assert rel_err(emb(xf), y.detach()) < 1e-6
        with pytest.raises(ValueError, match='same degree must be assigned'):
            emb.get_degrees_out(torch.arange(11))
        with pytest.raises(ValueError, match='duplicated indices'):
            type(emb)(4, 2, embedded_indices=[0, 0, 1, 2])
    else:
        with pytest.raises(ValueError, match='must be assigned to different feature indices'):
            type(emb)(6, list(emb.embedding_layers), [[0, 1, 2, 3], [3, 4]])


def test_sos_function_name_and_coefficients():
    from tfep_b200.nn.transformers import SOSPolynomialTransformerFunc, sos_polynomial_transformer
    assert SOSPolynomialTransformerFunc.apply is sos_polynomial_transformer
    par = cases.normal((5, 7, 3), 55)
    a0, c1, c2, c3 = SOSPolynomialTransformerFunc.get_sos_poly_coefficients(par)
    k0, k1 = par[:, 1::2], par[:, 2::2]
    assert torch.equal(a0, par[:, 0]) and torch.allclose(c1, (k0 ** 2).sum(1)) and torch.allclose(c2, (k0 * k1).sum(1))
    assert torch.allclose(c3, (k1 ** 2).sum(1) / 3) and c3.shape == (5, 3)


@pytest.mark.parametrize('axis,plane,expected', [(None, None, (0, 1)), (None, 0, (1, 0)), (0, None, (0, 1)), (2, None, (2, 0)),
                                                 (None, 3, (0, 3)), (1, 2, (1, 2))])
def test_oriented_flow_point_selection_and_partial_output(axis, plane, expected):
    """Automatic choice of the axis / plane points and ``return_partial`` (reference tests/nn/flows/test_oriented.py:
    test_automatic_axis_plane_selection, test_return_partial)."""
    from helpers import OracleFlowModule
    from tfep_b200.nn.flows import OrientedFlow
    inner, _ = cases.build_oracle(cases.wrapper_cases()['oriented']['inner'], torch.float32)
    flow = OrientedFlow(OracleFlowModule(inner), axis_point_idx=axis, plane_point_idx=plane)
    assert (int(flow._axis_point_idx), int(flow._plane_point_idx)) == expected
    x = cases.normal((3, 12), 5)
    y, ld = flow(x)
    assert y.shape == (3, 12) and ld.shape == (3,)
    flow.return_partial = True
    y, ld = flow(x)
    assert y.shape == (3, 9)
