"""Transformer fused into the output-layer product of the tensor-core conditioner (``precision='bf16'``; tfepb_tc_tx).

For a MAF whose transformer is ONE elementary kernel of kind affine / SOS with two polynomials / Moebius on 3-vectors /
neural spline with 8 bins over all features, the reference's ``parameters = conditioner(x); y, log_det = transformer(x, parameters)``
(nn/flows/autoregressive.py:144-177) runs as: hidden layers on tfepb_tc_gemm, then the output-layer product whose
epilogue applies the transformer to the accumulators -- the (batch x n_parameters) matrix, the largest tensor of the
layer, is neither written nor read, forward or backward (``_ops.MadeTxFunctionTC``).

The epilogue wants the parameters of whole units inside 16-column chunks (8 affine features x 2, 3 SOS features x 5 + 1 pad
column, 5 Moebius vectors x 3 + 1 pad column; a spline feature takes the 32 columns of a sub-tile).  This module derives that layout from the degree-sorted feature-major packing
of ``MAF._pack`` (units stay in degree order, so the staircase ranges survive): the row index map of the padded output
layer, the x / y columns of every unit and the k-block / row ranges of the padded layer.
"""

import torch

from . import _ops
from ._pack import MadePlan


def eligibility(maf, pk):
    """None if the fused epilogue covers the layer, else the reason (a string)."""
    if pk is False:
        return 'the transformer is not a native program'
    if maf._n_conditioner_indices > 0 or maf.has_fixed_indices:
        return 'conditioning / fixed features'
    if maf._embedding is not None:
        return 'the conditioner input goes through an embedding'
    parts = pk['parts']
    if len(parts) != 1:
        return 'mixed transformer'
    part = parts[0]
    if part.n_features != len(maf._degrees_in_host) or sorted(part.x_columns().tolist()) != list(range(part.n_features)):
        return 'the transformer does not map every feature'
    if part.kind == 'affine':
        pass
    elif part.kind == 'sos':
        if part.spec.n_polynomials != 2:
            return 'SOS transformer with more than two polynomials'
    elif part.kind == 'moebius':
        if part.spec.dimension != 3 or part.spec.unit_sphere not in (0, 1):
            return 'Moebius transformer other than 3-vectors (plain variants)'
        base = pk['bases'][0].long()
        b3 = base.view(-1, 3)
        if not (torch.equal(b3[:, 1], b3[:, 0] + 1) and torch.equal(b3[:, 2], b3[:, 0] + 2) and bool((b3[:, 0] % 3 == 0).all())):
            return 'the features of a Moebius vector are not consecutive in degree order'
    elif part.kind == 'spline':
        if part.spec.n_bins_int != 8:
            return 'neural spline with a number of bins other than 8'
    else:
        return f'transformer kind {part.kind!r}'
    if pk['plan'].n_layers < 2:
        return 'conditioner without hidden layers'
    return None


class TcTxPlan:
    """Padded chunk layout of the output layer of ``maf`` and the tables of the fused epilogue."""

    def __init__(self, maf, pk):
        part = pk['parts'][0]
        plan = pk['plan']
        self.kind = part.kind
        if self.kind == 'spline':
            # one feature per 32-column sub-tile: its 23..27 parameters, then padding
            upc, ppu, width = 1, part.n_params, 32
        else:
            upc, ppu, width = _ops.TCTX_UNITS_PER_CHUNK[self.kind], _ops.TCTX_COLUMNS_PER_UNIT[self.kind], 16
        n_rows = part.n_features * part.n_params                 # rows of the packed output layer
        assert n_rows == len(plan.perms[-1])
        n_units = n_rows // ppu
        n_chunks = (n_units + upc - 1) // upc
        self.n_padded = n_chunks * width
        # padded position -> packed row (-1 = zero row): units are consecutive runs of `ppu` packed rows
        t = torch.arange(self.n_padded)
        within, chunk = t % width, t // width
        src = chunk * (upc * ppu) + within
        valid = (within < upc * ppu) & (src < n_rows)
        out_order = torch.where(valid, plan.perms[-1][src.clamp(max=n_rows - 1)], torch.full_like(src, -1))
        #: the degree-sorted plan of the conditioner with its output layer in the padded chunk layout
        self.plan = MadePlan(maf._conditioner._degree_chain, out_order=out_order)
        # x / y columns of every unit, in unit (= packed) order
        base = pk['bases'][0].long()
        xcols = part.x_columns().long()
        if self.kind == 'moebius':
            order = torch.argsort(base.view(-1, 3)[:, 0])
            self.cols = xcols.view(-1, 3)[order].reshape(-1).to(torch.int32)
        else:
            order = torch.argsort(base)
            self.cols = xcols[order].to(torch.int32)
        assert torch.equal(base[order] if self.kind != 'moebius' else base.view(-1, 3)[order, 0],
                           torch.arange(n_units) * ppu)
        self.max_radius = float(getattr(part.spec, 'max_radius', 0.0))
        self.unit_sphere = int(getattr(part.spec, 'unit_sphere', 0))
        self._order = order                                      # unit -> local feature of the transformer
        self._spec = part.spec
        self._dev = {}

    def tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = dict(kind=self.kind, cols=self.cols.to(device), max_radius=self.max_radius,
                                  unit_sphere=self.unit_sphere)
        spec = self._dev[key]
        if self.kind == 'spline':
            # domain of every unit in unit order (re-derived when the module's buffers change), options of spline.py:166-182
            sp = self._spec
            dom = sp.domain_tensors(torch.float32, device)
            tag = tuple((t.data_ptr(), t._version) for t in dom)
            if spec.get('_dom_tag') != tag:
                order = self._order.to(device)
                x0, xf, y0, yf = (t.index_select(0, order).contiguous() for t in dom)
                spec['spline'] = dict(x0=x0, xf=xf, y0=y0, yf=yf, min_bin_size=float(sp.min_bin_size), min_slope=float(sp.min_slope),
                                      flags=int(sp.circular) | int(sp.identity_slopes) << 1 | int(sp.learn_lower) << 2 |
                                      int(sp.learn_upper) << 3)
                spec['_dom_tag'] = tag
        return spec

    def forward(self, maf, pk, x):
        kb_fwd, kb_bwd, rr_w = self.plan.tc_ranges(x.device)
        pw, pb = maf._conditioner.packed_weights(self.plan)
        return _ops.made_tx_forward_tc(x, list(pw), list(pb), kb_fwd, kb_bwd, rr_w, self.tables(x.device))
