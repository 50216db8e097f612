"""Seeded synthetic workloads shared by the golden generator, the tests and bench.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing here reads the
reference; everything is reproducible from integer seeds with CPU
``torch.Generator`` streams (identical across hosts for one torch version; the
golden files carry checksums so that drift would be detected).
"""

import math

import torch

from . import flow_oracle as fo


def seeded_state(masks, seed, dtype=torch.float32, weight_norm=True, gain=1.0, prefix=''):
    """Random conditioner parameters with the key names of the reference ``state_dict``.

    Distribution of nn.Linear's default init (U(+-1/sqrt(fan_in)), as used by
    nn/masked.py:164-176), masked, then g = ||v||_row as in nn/masked.py:391-392,
    optionally scaled by ``gain`` to make the transformer parameters more varied.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for li, mask in enumerate(masks):
        out_f, in_f = mask.shape
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        w = (w * mask.double()).to(dtype)
        key = f'{prefix}_conditioner.layers.{2 * li}.'
        if weight_norm:
            sd[key + 'weight_v'] = w
            sd[key + 'weight_g'] = (torch.linalg.norm(w.double(), dim=1, keepdim=True) * gain).to(dtype)
        else:
            sd[key + 'weight'] = w * gain
        sd[key + 'bias'] = b.to(dtype)
    return sd


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def uniform(shape, seed, lo, hi, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)


def normal(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)


# ---------------------------------------------------------------------------
# transformer-level cases: name -> (spec factory, n_features, x generator)
# ---------------------------------------------------------------------------

def _spline(n, dtype, **kw):
    x0 = kw.pop('x0', -2.0)
    xf = kw.pop('xf', 2.0)
    return fo.Spline(x0=torch.full((n,), x0, dtype=dtype), xf=torch.full((n,), xf, dtype=dtype), **kw)


def as_double(spec):
    """The same transformer spec with float64 constants (to evaluate fp32 inputs in double)."""
    if isinstance(spec, fo.Spline):
        return fo.Spline(x0=spec.x0.double(), xf=spec.xf.double(), n_bins=spec.n_bins, y0=spec.y0.double(),
                         yf=spec.yf.double(), circular=spec.circular,
                         identity_boundary_slopes=spec.identity_boundary_slopes,
                         learn_lower_bound=spec.learn_lower_bound, learn_upper_bound=spec.learn_upper_bound,
                         min_bin_size=spec.min_bin_size, min_slope=spec.min_slope)
    if isinstance(spec, fo.Mixed):
        return fo.Mixed([as_double(t) for t in spec.transformers], [i.tolist() for i in spec.indices])
    if isinstance(spec, fo.Shift) and spec.periodic_limits is not None:
        return fo.Shift(spec.periodic_indices, spec.periodic_limits.double())
    return spec


def double_reference(spec, x, par, inverse=False):
    """Evaluate ``spec`` in float64 on (the exact values of) lower-precision inputs."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        d = as_double(spec)
        return (d.inverse if inverse else d.forward)(x.double(), par.double())
    finally:
        torch.set_default_dtype(old)


def transformer_cases(dtype=torch.float32):
    """name -> (spec, n_features, x, params); small and cheap."""
    B = 24
    cases = {}

    def add(name, spec, n, x, pscale=1.0, seed=11):
        par = normal((B, len(spec.identity_params(n))), seed, dtype) * pscale
        cases[name] = (spec, n, x, par)

    add('affine', fo.Affine(), 7, normal((B, 7), 1, dtype))
    add('shift', fo.Shift(), 5, normal((B, 5), 17, dtype))
    add('shift_periodic', fo.Shift(periodic_indices=torch.tensor([0, 3]),
                                   periodic_limits=torch.tensor([-math.pi, math.pi], dtype=dtype)), 5,
        normal((B, 5), 18, dtype) * 2.0, pscale=2.0)
    add('spline_k8', _spline(6, dtype, n_bins=8), 6, normal((B, 6), 2, dtype) * 1.5)      # some in the tails
    add('spline_k5_circular', _spline(5, dtype, n_bins=5, circular=True, x0=-math.pi, xf=math.pi), 5,
        uniform((B, 5), 3, -math.pi, math.pi, dtype) * 0.999, pscale=1.5)
    add('spline_k4_idslopes', _spline(4, dtype, n_bins=4, identity_boundary_slopes=True), 4, normal((B, 4), 4, dtype))
    add('spline_k4_circ_idslopes', _spline(4, dtype, n_bins=4, circular=True, identity_boundary_slopes=True), 4,
        uniform((B, 4), 5, -2.0, 2.0, dtype) * 0.999)
    add('spline_k6_learn_lower', _spline(3, dtype, n_bins=6, learn_lower_bound=True), 3, normal((B, 3), 6, dtype),
        pscale=0.5)
    add('spline_k6_learn_upper', _spline(3, dtype, n_bins=6, learn_upper_bound=True), 3, normal((B, 3), 7, dtype),
        pscale=0.5)
    add('spline_k3_learn_both', _spline(3, dtype, n_bins=3, learn_lower_bound=True, learn_upper_bound=True), 3,
        normal((B, 3), 8, dtype), pscale=0.5)
    add('sos2', fo.SOS(2), 5, normal((B, 5), 9, dtype))
    add('sos3', fo.SOS(3), 5, normal((B, 5), 10, dtype))
    add('moebius_d3', fo.Moebius(dimension=3), 6, normal((B, 6), 12, dtype))
    add('moebius_d2', fo.Moebius(dimension=2), 6, normal((B, 6), 13, dtype))
    add('moebius_d4_unit', fo.Moebius(dimension=4, unit_sphere=True), 8,
        torch.nn.functional.normalize(normal((B, 2, 4), 14, dtype), dim=-1).reshape(B, 8))
    add('symmoebius_d3', fo.SymMoebius(dimension=3), 6, normal((B, 6), 19, dtype))
    add('symmoebius_d2', fo.SymMoebius(dimension=2, max_radius=0.9), 6, normal((B, 6), 20, dtype), pscale=2.0)
    mixed = fo.Mixed([_spline(2, dtype, n_bins=4, circular=True, x0=-math.pi, xf=math.pi),
                      fo.Affine(), _spline(3, dtype, n_bins=4, x0=-4.0, xf=4.0)],
                     [[1, 4], [0, 6], [2, 3, 5]])
    xm = normal((B, 7), 15, dtype)
    xm[:, [1, 4]] = uniform((B, 2), 16, -math.pi, math.pi, dtype) * 0.999
    add('mixed', mixed, 7, xm)
    return cases


# ---------------------------------------------------------------------------
# MAF-level cases
# ---------------------------------------------------------------------------

def maf_cases(dtype=torch.float32):
    """name -> dict(degrees_in, spec, hidden_layers, weight_norm, x, seed[, invertible])."""
    B = 16
    c = {}

    def add(name, degrees_in, spec, x, hidden_layers=2, weight_norm=True, seed=100, invertible=True, gain=2.0,
            embedding=None):
        c[name] = dict(degrees_in=torch.as_tensor(degrees_in), spec=spec, hidden_layers=hidden_layers,
                       weight_norm=weight_norm, x=x, seed=seed, invertible=invertible, gain=gain, embedding=embedding)

    add('affine_asc', fo.gen_degrees(8), fo.Affine(), normal((B, 8), 21, dtype))
    add('affine_desc_nown', fo.gen_degrees(8, order='descending'), fo.Affine(), normal((B, 8), 22, dtype),
        weight_norm=False)
    add('affine_cond_h1', fo.gen_degrees(9, conditioning_indices=[0, 4]), fo.Affine(), normal((B, 9), 23, dtype),
        hidden_layers=1)
    add('affine_h4', fo.gen_degrees(6, order='descending'), fo.Affine(), normal((B, 6), 24, dtype), hidden_layers=4)
    add('spline_circ', fo.gen_degrees(6), _spline(6, dtype, n_bins=8, circular=True, x0=-math.pi, xf=math.pi),
        uniform((B, 6), 25, -math.pi, math.pi, dtype) * 0.999)
    add('spline_desc_cond', fo.gen_degrees(7, order='descending', conditioning_indices=[2]),
        _spline(6, dtype, n_bins=5, x0=-4.0, xf=4.0), normal((B, 7), 26, dtype))
    add('sos2', fo.gen_degrees(6), fo.SOS(2), normal((B, 6), 27, dtype), invertible=False, gain=1.0)
    add('moebius_d3', fo.gen_degrees(6, repeats=3), fo.Moebius(dimension=3), normal((B, 6), 28, dtype))
    add('symmoebius_d3', fo.gen_degrees(6, repeats=3, order='descending'), fo.SymMoebius(dimension=3),
        normal((B, 6), 37, dtype))
    add('moebius_d2_cond', fo.gen_degrees(8, conditioning_indices=[0, 1], repeats=2, order='descending'),
        fo.Moebius(dimension=2), normal((B, 8), 29, dtype))
    mixed = fo.Mixed([_spline(2, dtype, n_bins=4, circular=True, x0=-math.pi, xf=math.pi),
                      _spline(4, dtype, n_bins=4, x0=-4.0, xf=4.0)], [[2, 5], [0, 1, 3, 4]])
    xm = normal((B, 6), 30, dtype)
    xm[:, [2, 5]] = uniform((B, 2), 31, -math.pi, math.pi, dtype) * 0.999
    add('mixed_splines', fo.gen_degrees(6), mixed, xm)
    add('repeated_degrees', torch.tensor([0, 0, 1, 2, 2, 2, 1]), fo.Affine(), normal((B, 7), 32, dtype))
    # PeriodicEmbedding of the conditioner input (what the reference's MixedMAFMap runs, app/mixedmaf.py:341-353)
    xe = normal((B, 7), 34, dtype)
    xe[:, [1, 4, 5]] = uniform((B, 3), 35, 0.0, 1.0, dtype)
    add('spline_embed_periodic', fo.gen_degrees(7, conditioning_indices=[2]),
        fo.Mixed([_spline(3, dtype, n_bins=4, circular=True, x0=0.0, xf=1.0), _spline(3, dtype, n_bins=4, x0=-4.0, xf=4.0)],
                 [[0, 3, 4], [1, 2, 5]]),                 # indices among the 6 mapped features: columns 1, 4, 5 are circular
        xe, embedding=fo.PeriodicEmbed(7, [0.0, 1.0], [1, 4, 5]))
    add('affine_embed_all_desc', fo.gen_degrees(5, order='descending'), fo.Affine(), normal((B, 5), 36, dtype),
        hidden_layers=1, embedding=fo.PeriodicEmbed(5, torch.tensor([-math.pi, math.pi], dtype=dtype)))
    add('shift_periodic_cond', fo.gen_degrees(7, conditioning_indices=[3]),
        fo.Shift(periodic_indices=torch.tensor([1, 4]), periodic_limits=torch.tensor([0.0, 2.0], dtype=dtype)),
        normal((B, 7), 33, dtype))
    return c


def wrapper_cases(dtype=torch.float32):
    """name -> dict(layers=[(kind, kwargs), ...] outermost first, inner=<MAF case on the propagated features>, x,
    invertible).  Kinds: 'partial', 'centroid', 'oriented' (reference nn/flows/{partial,centroid,oriented}.py)."""
    B = 12
    c = {}

    def inner(n, seed, spec=None):
        return dict(degrees_in=fo.gen_degrees(n), spec=spec or fo.Affine(), hidden_layers=2, weight_norm=True, seed=seed,
                    gain=2.0, embedding=None)

    def add(name, layers, n_in, n_inner, seed, invertible=True, spec=None):
        c[name] = dict(layers=layers, inner=inner(n_inner, 300 + seed, spec), x=normal((B, n_in), 200 + seed, dtype) * 1.5,
                       invertible=invertible)

    add('partial', [('partial', dict(fixed_indices=[1, 4]))], 7, 5, 1)
    add('partial_return', [('partial', dict(fixed_indices=[0], return_partial=True))], 5, 4, 2, invertible=False)
    add('centroid', [('centroid', dict(space_dimension=3))], 12, 9, 3)
    add('centroid_weighted_subset',
        [('centroid', dict(space_dimension=3, subset_point_indices=[0, 2, 3], weights=[1.0, 2.0, 3.0], fixed_point_idx=1,
                           origin=[0.5, -1.0, 2.0]))], 12, 9, 4)
    add('centroid_2d_no_back', [('centroid', dict(space_dimension=2, translate_back=False))], 8, 6, 5, invertible=False)
    add('oriented', [('oriented', dict())], 12, 9, 6, spec=_spline(9, dtype, n_bins=5, x0=-6.0, xf=6.0))
    add('oriented_z_yz', [('oriented', dict(axis_point_idx=2, plane_point_idx=0, axis='z', plane='yz'))], 12, 9, 7)
    add('centroid_oriented', [('centroid', dict(space_dimension=3, fixed_point_idx=1)),
                              ('oriented', dict(axis_point_idx=0, plane_point_idx=2))], 15, 9, 8)
    return c


def build_wrapper_oracle(case, dtype=torch.float32):
    """(oracle flow, state dict of the innermost MAF) of a wrapper case."""
    from . import wrappers_oracle as wo
    flow, sd = build_oracle(case['inner'], dtype)
    for kind, kw in reversed(case['layers']):
        kw = {k: (torch.tensor(v, dtype=dtype) if k in ('weights', 'origin') else v) for k, v in kw.items()}
        flow = {'partial': wo.Partial, 'centroid': wo.Centroid, 'oriented': wo.Oriented}[kind](flow, **kw)
    return flow, sd


def embedding_cases():
    """name -> (builder(namespace) -> module, n_features_in, degrees_in): embeddings with learnable parameters, built from
    a namespace that provides FlipInvariantEmbedding / MixedEmbedding / PeriodicEmbedding (the reference or tfep_b200)."""
    def flip(ns):
        return ns.FlipInvariantEmbedding(n_features_in=11, embedding_dimension=5, embedded_indices=[1, 2, 3, 4, 7, 8, 9, 10])

    def mixed_flips(ns):
        return ns.MixedEmbedding(12, [ns.FlipInvariantEmbedding(4, 3, vector_dimension=4, hidden_layer_width=8),
                                      ns.FlipInvariantEmbedding(6, 2, vector_dimension=2, hidden_layer_width=6)],
                                 [[0, 1, 2, 3], [5, 6, 8, 9, 10, 11]])

    def mixed_periodic(ns):
        return ns.MixedEmbedding(9, [ns.PeriodicEmbedding(3, [-math.pi, math.pi]),
                                     ns.FlipInvariantEmbedding(4, 6, hidden_layer_width=16)], [[0, 4, 8], [2, 3, 5, 6]])

    deg11 = torch.tensor([0, 1, 1, 1, 1, 2, 3, 4, 4, 4, 4])
    deg12 = torch.tensor([0, 0, 0, 0, 1, 2, 2, 3, 4, 4, 5, 5])
    deg9 = torch.tensor([0, 1, 2, 2, 3, 2, 2, 4, 5])
    return {'flip': (flip, 11, deg11), 'mixed_flips': (mixed_flips, 12, deg12), 'mixed_periodic': (mixed_periodic, 9, deg9)}


def logger_script(logger_cls, directory):
    """One fixed sequence of TFEPLogger operations (reference tfep/io/log.py) on a fresh ``directory``; returns the
    logger and a dict of everything that was read back.  Used to generate the golden log directory with the reference
    and to replay it with tfep_b200.io.TFEPLogger."""
    import warnings

    class _Loader:                                  # the three attributes the logger reads from a DataLoader
        batch_size, drop_last, dataset = 4, False, range(10)

    def batch(first, n, seed):
        idx = torch.arange(first, first + n)
        pot = normal((n,), seed)
        pot[n // 2] = float('nan') if seed % 2 == 0 else pot[n // 2]
        return {'dataset_sample_index': idx, 'trajectory_sample_index': 3 * idx + 1, 'potential': pot,
                'log_det_J': normal((n,), seed + 100).double()}

    log = logger_cls(save_dir_path=directory, data_loader=_Loader())
    reads = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        # training: epoch 0 complete (3 batches: 4 + 4 + 2 samples), epoch 1 only its middle batch
        for b, (first, n) in enumerate(((0, 4), (4, 4), (8, 2))):
            log.save_train_tensors(batch(first, n, 10 + b), epoch_idx=0, batch_idx=b)
        log.save_train_tensors(batch(4, 4, 20), step_idx=4)
        log.save_train_tensors({'loss_weight': normal((4,), 21)}, epoch_idx=1, batch_idx=1)      # no indices: warns
        # evaluation at step 7: two batches appended, then an update that overwrites two samples and adds one
        log.save_eval_tensors(batch(6, 3, 30), step_idx=7)
        log.save_eval_tensors(batch(0, 3, 31), epoch_idx=2, batch_idx=1)
        upd = batch(1, 3, 32)
        upd['dataset_sample_index'] = torch.tensor([2, 7, 9])
        upd['trajectory_sample_index'] = 3 * upd['dataset_sample_index'] + 1
        log.save_eval_tensors(upd, step_idx=7, update=True)
    reads['train_e0'] = log.read_train_tensors(epoch_idx=0, as_numpy=True)
    reads['train_e0_b1'] = log.read_train_tensors(names=['potential'], step_idx=1, as_numpy=True)
    reads['train_e0_nonan'] = log.read_train_tensors(epoch_idx=0, remove_nans=True, as_numpy=True)
    reads['train_e1'] = log.read_train_tensors(epoch_idx=1, as_numpy=True)
    reads['train_e1_nonan_pot'] = log.read_train_tensors(epoch_idx=1, remove_nans='potential', as_numpy=True)
    reads['eval_s7'] = log.read_eval_tensors(step_idx=7, as_numpy=True)
    reads['eval_s7_nonan'] = log.read_eval_tensors(names=['log_det_J'], step_idx=7, remove_nans='potential', as_numpy=True)
    reads['eval_s7_sorted'] = log.read_eval_tensors(step_idx=7, sort_by='trajectory_sample_index', as_numpy=True)
    return log, reads


def build_oracle(case, dtype=torch.float32):
    m = fo.MafOracle(case['degrees_in'], case['spec'], hidden_layers=case['hidden_layers'],
                     weight_norm=case['weight_norm'], embedding=case.get('embedding'))
    sd = seeded_state([k.to(dtype) for k in m.masks], case['seed'], dtype, case['weight_norm'], case['gain'])
    return m.load(sd), sd


# ---------------------------------------------------------------------------
# BASELINE.json configurations
# ---------------------------------------------------------------------------

def cfg_flow(name, dtype=torch.float32, n_layers=None, D=None, hidden_layers=2):
    """Return a list of (MafOracle, state_dict) for a BASELINE.json configuration."""
    if name == 'cfg1':           # 2 x MAF Affine, D=66
        D, L = D or 66, n_layers or 2
        mk = lambda: fo.Affine()
    elif name == 'cfg2':         # 4 x MAF circular spline K=8, D=66
        D, L = D or 66, n_layers or 4
        mk = lambda: fo.Spline(x0=torch.full((D,), -math.pi, dtype=dtype), xf=torch.full((D,), math.pi, dtype=dtype),
                               n_bins=8, circular=True)
    elif name in ('cfg2cond', 'cfg2condemb'):     # cfg2 with the first two atoms (6 features) as conditioning features (degree -1):
        D, L = D or 66, n_layers or 4            # they steer the conditioner and pass through unchanged
        mk = lambda: fo.Spline(x0=torch.full((D - 6,), -math.pi, dtype=dtype), xf=torch.full((D - 6,), math.pi, dtype=dtype),
                               n_bins=8, circular=True)
    elif name in ('cfg2mix', 'cfg2mixemb'):      # cfg2's "dihedral / Cartesian mix": circular splines on the torsions
        D, L = D or 66, n_layers or 4            # f % 3 == 2, ordinary splines on [-4, 4] for the others;
                                                 # 'emb': torsions enter the conditioner as (cos, sin), app/mixedmaf.py:341-353
        tors = [f for f in range(D) if f % 3 == 2]
        cart = [f for f in range(D) if f % 3 != 2]
        mk = lambda: fo.Mixed([fo.Spline(x0=torch.full((len(tors),), -math.pi, dtype=dtype),
                                         xf=torch.full((len(tors),), math.pi, dtype=dtype), n_bins=8, circular=True),
                               fo.Spline(x0=torch.full((len(cart),), -4.0, dtype=dtype),
                                         xf=torch.full((len(cart),), 4.0, dtype=dtype), n_bins=8)], [tors, cart])
    elif name == 'cfg5':         # 8 x MAF spline K=8 (non circular), D=3000 (scaled down in tests)
        D, L = D or 3000, n_layers or 8
        mk = lambda: fo.Spline(x0=torch.full((D,), -5.0, dtype=dtype), xf=torch.full((D,), 5.0, dtype=dtype), n_bins=8)
    elif name == 'cfg3':         # 3 x SOS(2) + 3 x Moebius(d=3), D=300
        D, L = D or 300, n_layers or 6
        mk = None
    else:
        raise ValueError(name)
    flows = []
    for l in range(L):
        order = 'ascending' if l % 2 == 0 else 'descending'
        if name == 'cfg3':
            if l < L // 2:
                spec, deg = fo.SOS(2), fo.gen_degrees(D, order=order)
            else:
                spec, deg = fo.Moebius(dimension=3), fo.gen_degrees(D, order=order, repeats=3)
        elif name in ('cfg2cond', 'cfg2condemb'):
            spec, deg = mk(), fo.gen_degrees(D, order=order, conditioning_indices=list(range(6)))
        else:
            spec, deg = mk(), fo.gen_degrees(D, order=order)
        emb = None
        if name == 'cfg2mixemb':
            emb = fo.PeriodicEmbed(D, torch.tensor([-math.pi, math.pi], dtype=dtype), [f for f in range(D) if f % 3 == 2])
        if name == 'cfg2condemb':       # every feature, the conditioning ones included, enters as (cos, sin)
            emb = fo.PeriodicEmbed(D, torch.tensor([-math.pi, math.pi], dtype=dtype))
        m = fo.MafOracle(deg, spec, hidden_layers=hidden_layers, embedding=emb)
        sd = seeded_state([k.to(dtype) for k in m.masks], 1234 + l, dtype)
        flows.append((m.load(sd), sd))
    return flows


def cfg_input(name, batch, dtype=torch.float32, D=None):
    if name in ('cfg2', 'cfg2cond', 'cfg2condemb'):
        return uniform((batch, D or 66), 0, -math.pi, math.pi, dtype) * 0.999
    if name in ('cfg2mix', 'cfg2mixemb'):        # torsions uniform in (-pi, pi), Cartesians normal with a few samples in the spline tails
        D = D or 66
        x = normal((batch, D), 0, dtype) * 1.6
        tors = [f for f in range(D) if f % 3 == 2]
        x[:, tors] = uniform((batch, len(tors)), 1, -math.pi, math.pi, dtype) * 0.999
        return x
    return normal((batch, D or {'cfg1': 66, 'cfg3': 300, 'cfg5': 3000}[name]), 0, dtype)
