"""Shared test helpers: build tfep_b200 modules from the oracle's case descriptions."""

import os

import numpy as np
import torch

from oracle import cases
from oracle import flow_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def to_module(spec):
    """oracle transformer spec -> tfep_b200 transformer module."""
    from tfep_b200.nn import transformers as T
    if isinstance(spec, fo.Affine):
        return T.AffineTransformer()
    if isinstance(spec, fo.Shift):
        return T.VolumePreservingShiftTransformer(spec.periodic_indices, spec.periodic_limits)
    if isinstance(spec, fo.Spline):
        return T.NeuralSplineTransformer(
            x0=spec.x0.clone(), xf=spec.xf.clone(), n_bins=spec.n_bins, y0=spec.y0.clone(), yf=spec.yf.clone(),
            circular=spec.circular, identity_boundary_slopes=spec.identity_boundary_slopes,
            learn_lower_bound=spec.learn_lower_bound, learn_upper_bound=spec.learn_upper_bound,
            min_bin_size=spec.min_bin_size, min_slope=spec.min_slope)
    if isinstance(spec, fo.SOS):
        return T.SOSPolynomialTransformer(spec.n_polynomials)
    if isinstance(spec, fo.Moebius):
        return T.MoebiusTransformer(spec.dimension, max_radius=spec.max_radius, unit_sphere=spec.unit_sphere)
    if isinstance(spec, fo.SymMoebius):
        return T.SymmetrizedMoebiusTransformer(spec.dimension, max_radius=spec.max_radius, identity_eps=spec.identity_eps)
    if isinstance(spec, fo.Mixed):
        return T.MixedTransformer([to_module(t) for t in spec.transformers], [i.tolist() for i in spec.indices])
    raise TypeError(spec)


def to_maf(case, state_dict, device=None, dtype=None):
    """oracle MAF case (+ its seeded state) -> tfep_b200.nn.flows.MAF with the same parameters."""
    from tfep_b200.nn.embeddings import PeriodicEmbedding
    from tfep_b200.nn.flows import MAF
    emb = case.get('embedding')
    if emb is not None:
        emb = PeriodicEmbedding(emb.n_features_in, emb.limits, emb.periodic_indices)
    maf = MAF(degrees_in=case['degrees_in'], transformer=to_module(case['spec']), hidden_layers=case['hidden_layers'],
              embedding=emb, weight_norm=case['weight_norm'], initialize_identity=False)
    if dtype is not None:
        maf = maf.to(dtype)
    missing, unexpected = maf.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if 'weight' in k or 'bias' in k], missing
    if device is not None:
        maf = maf.to(device)
    return maf


def to_wrapper(case, inner_flow, dtype=torch.float32):
    """The tfep_b200 wrapper flows of a wrapper case (oracle.cases.wrapper_cases) around ``inner_flow``."""
    from tfep_b200.nn.flows import CenteredCentroidFlow, OrientedFlow, PartialFlow
    flow = inner_flow
    for kind, kw in reversed(case['layers']):
        kw = {k: (torch.tensor(v, dtype=dtype) if k in ('weights', 'origin') else v) for k, v in kw.items()}
        flow = {'partial': PartialFlow, 'centroid': CenteredCentroidFlow, 'oriented': OrientedFlow}[kind](flow, **kw)
    return flow


class OracleFlowModule(torch.nn.Module):
    """A CPU flow for host-side tests of the wrapper logic: forward / inverse of an oracle flow (autograd works)."""

    def __init__(self, oracle):
        super().__init__()
        self.oracle = oracle

    def forward(self, x):
        return self.oracle.forward(x)

    def inverse(self, y):
        return self.oracle.inverse(y)

    def n_parameters(self):
        return 0


def cfg_flow_modules(name, device, n_layers=None, D=None, dtype=torch.float32, hidden_layers=2):
    """BASELINE.json configuration as (tfep_b200 SequentialFlow on `device`, [oracle flows])."""
    from tfep_b200.nn.flows import SequentialFlow
    flows = cases.cfg_flow(name, torch.float32, n_layers=n_layers, D=D, hidden_layers=hidden_layers)
    mafs = []
    for m, sd in flows:
        case = dict(degrees_in=m.degrees_in, spec=m.transformer, hidden_layers=hidden_layers, weight_norm=True,
                    embedding=m.embedding)
        mafs.append(to_maf(case, {k: v.to(dtype) for k, v in sd.items()}, dtype=dtype))
    return SequentialFlow(*mafs).to(device), flows


def rel_err(a, b):
    """max |a - b| / (1 + |b|) over all elements, in double."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float(((a - b).abs() / (1 + b.abs())).max()) if a.numel() else 0.0
