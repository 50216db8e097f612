"""Transformer fused into the output-layer product of the tensor-core conditioner (``precision='bf16'``; tfepb_tc_tx).

For a MAF whose transformer is ONE elementary kernel of kind affine / SOS with two polynomials / Moebius on 3-vectors /
neural spline with 8 bins over all features -- or a MixedTransformer whose children are all 8-bin splines --, the reference's ``parameters = conditioner(x); y, log_det = transformer(x, parameters)``
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
    if maf._n_conditioner_indices > 0:
        return 'the conditioner reads a subset of the features'
    parts = pk['parts']
    columns = sorted(c for p in parts for c in p.x_columns().tolist())
    if columns != sorted(maf._mapped_host.tolist()):
        return 'the transformer program does not cover the mapped features'
    if len(parts) > 1:
        # a MixedTransformer is covered when all its children are 8-bin splines (any options: the table is per unit)
        if not all(p.kind == 'spline' and p.spec.n_bins_int == 8 for p in parts):
            return 'mixed transformer with children other than 8-bin neural splines'
    else:
        part = parts[0]
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
        parts = pk['parts']
        plan = pk['plan']
        self.kind = parts[0].kind
        n_rows = len(plan.perms[-1])                              # rows of the packed output layer
        if self.kind == 'spline':
            # one feature per 32-column sub-tile: its 23..27 parameters, then padding.  Units = the features of all parts in
            # packed (degree) order; the packed rows of unit u start at its base and run for its part's parameter count.
            units = sorted((int(pk['bases'][pi][f]), pi, f) for pi, p in enumerate(parts) for f in range(p.n_features))
            starts = torch.tensor([u[0] for u in units])
            counts = torch.tensor([parts[u[1]].n_params for u in units])
            assert int(starts[0]) == 0 and torch.equal(starts[1:], (starts + counts)[:-1]) and int((starts + counts)[-1]) == n_rows
            n_units = len(units)
            self.n_padded = n_units * 32
            t = torch.arange(self.n_padded)
            within, unit = t % 32, t // 32
            valid = within < counts[unit]
            src = starts[unit] + within
            self.cols = torch.tensor([int(parts[pi].x_columns()[f]) for _, pi, f in units], dtype=torch.int32)
            self._units = [(pi, f) for _, pi, f in units]
        else:
            part = parts[0]
            upc, ppu = _ops.TCTX_UNITS_PER_CHUNK[self.kind], _ops.TCTX_COLUMNS_PER_UNIT[self.kind]
            assert n_rows == part.n_features * part.n_params
            n_units = n_rows // ppu
            n_chunks = (n_units + upc - 1) // upc
            self.n_padded = n_chunks * 16
            # padded position -> packed row (-1 = zero row): units are consecutive runs of `ppu` packed rows
            t = torch.arange(self.n_padded)
            within, chunk = t % 16, t // 16
            src = chunk * (upc * ppu) + within
            valid = (within < upc * ppu) & (src < n_rows)
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
        out_order = torch.where(valid, plan.perms[-1][src.clamp(max=n_rows - 1)], torch.full_like(src, -1))
        #: the degree-sorted plan of the conditioner with its output layer in the padded chunk layout
        self.plan = MadePlan(maf._conditioner._degree_chain, out_order=out_order)
        self.max_radius = float(getattr(parts[0].spec, 'max_radius', 0.0))
        self.unit_sphere = int(getattr(parts[0].spec, 'unit_sphere', 0))
        #: conditioning features (degree -1): read by the conditioner, copied through by the flow
        self.passthrough = bool(maf.has_fixed_indices)
        self._parts = parts
        self._dev = {}

    def tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = dict(kind=self.kind, cols=self.cols.to(device), max_radius=self.max_radius,
                                  unit_sphere=self.unit_sphere, passthrough=self.passthrough)
        spec = self._dev[key]
        if self.kind == 'spline':
            # per unit: domain, minimum bin size / slope and option bits (re-derived when a child's domain buffers change)
            doms = [p.spec.domain_tensors(torch.float32, device) for p in self._parts]
            tag = tuple((t.data_ptr(), t._version) for dom in doms for t in dom)
            if spec.get('_dom_tag') != tag:
                host = [[t.detach().cpu() for t in dom] for dom in doms]
                table = torch.zeros(len(self._units), 8, dtype=torch.float32)
                flags = torch.zeros(len(self._units), dtype=torch.int32)
                for u, (pi, f) in enumerate(self._units):
                    sp = self._parts[pi].spec
                    table[u, 0], table[u, 1], table[u, 2], table[u, 3] = (host[pi][j][f] for j in range(4))
                    table[u, 4], table[u, 5] = float(sp.min_bin_size), float(sp.min_slope)
                    flags[u] = int(sp.circular) | int(sp.identity_slopes) << 1 | int(sp.learn_lower) << 2 | int(sp.learn_upper) << 3
                table[:, 6] = flags.view(torch.float32)
                spec['spline'] = table.to(device)
                spec['_dom_tag'] = tag
        return spec

    def forward(self, maf, pk, x):
        kb_fwd, kb_bwd, rr_w = self.plan.tc_ranges(x.device)
        pw, pb = maf._conditioner.packed_weights(self.plan)
        xc = None
        if maf._embedding is not None:
            # an embedding layer in front of the MADE (e.g. PeriodicEmbedding): the conditioner reads the embedded features, the
            # transformer maps x itself
            xc = maf._embedding(x)
            if xc.dtype != torch.float32:
                raise _ops._lib.TfepB200Error("precision='bf16' takes float32 inputs")
            xc = xc.contiguous()
        return _ops.made_tx_forward_tc(x, list(pw), list(pb), kb_fwd, kb_bwd, rr_w, self.tables(x.device), xc=xc)
