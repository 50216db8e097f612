"""Host side of the persistent inverse sweep (tfepb_maf_inverse_sweep, tfep_b200/csrc/maf_inverse.cu).

Builds, once per MAF layer and device, the tables the kernel walks: for every degree (ascending) the rows of
the packed output layer owned by the features of that degree, the hidden units that become computable after
them, and which transformer parts / features to invert (reference nn/flows/autoregressive.py:179-229 runs the
same dependency order as ``n_degrees`` full passes).  Integer work at construction time only.
"""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, dtype_code, stream_ptr

MAXL = 5
KIND = {'affine': 0, 'spline': 1, 'moebius': 2, 'shift': 3}

GROUP_DTYPE = np.dtype([('out_r0', '<i4'), ('out_r1', '<i4'), ('out_k', '<i4'), ('h_a', '<i4', (MAXL - 1,)),
                        ('h_b', '<i4', (MAXL - 1,)), ('h_k', '<i4', (MAXL - 1,)), ('part_first', '<i4'),
                        ('part_count', '<i4')])
PART_DTYPE = np.dtype([('kind', '<i4'), ('n_bins', '<i4'), ('circular', '<i4'), ('identity_boundary_slopes', '<i4'),
                       ('learn_lower_bound', '<i4'), ('learn_upper_bound', '<i4'), ('dimension', '<i4'),
                       ('unit_sphere', '<i4'), ('x0', '<u8'), ('xf', '<u8'), ('y0', '<u8'), ('yf', '<u8'),
                       ('min_bin_size', '<f8'), ('min_slope', '<f8'), ('max_radius', '<f8'), ('cols', '<u8'),
                       ('par_base', '<u8')])
GPART_DTYPE = np.dtype([('part', '<i4'), ('ids_offset', '<i4'), ('n_ids', '<i4')])


def eligibility(maf, pk):
    """None if the sweep kernel covers this MAF layer, else the reason."""
    if pk is False:
        return 'transformer does not lower to native parts'
    if pk['plan'].n_layers > MAXL:
        return 'too many linear layers'
    if maf._embedding is not None:
        from .nn.embeddings import PeriodicEmbedding
        if not isinstance(maf._embedding, PeriodicEmbedding):
            return 'only PeriodicEmbedding is applied inside the sweep kernel'
    for part in pk['parts']:
        if part.kind not in KIND:
            return f'no inverse for transformer kind {part.kind!r}'
        if part.kind == 'spline' and part.spec.n_bins_int > 32:
            return 'more than 32 spline bins'
        if part.kind == 'moebius' and part.spec.dimension > 16:
            return 'Moebius dimension above 16'
    return None


class SweepPlan:
    def __init__(self, maf, pk):
        why = eligibility(maf, pk)
        if why is not None:
            raise _lib.TfepB200Error(f'persistent inverse sweep unavailable: {why}')
        self.plan, self.parts = pk['plan'], pk['parts']
        plan = self.plan
        L = plan.n_layers
        self.L = L
        self.D = len(maf._degrees_in_host)
        self.E = len(plan.packed_degrees[0])           # width of the conditioner input (> D with an embedding)
        self.embedding = maf._embedding
        groups = np.zeros(0, dtype=GROUP_DTYPE)
        gparts, ids = [], []
        rows = []

        def hidden_rows(rec, degree):
            for l in range(1, L):
                a, b = plan.degree_rows(l, degree)
                rec['h_a'][l - 1], rec['h_b'][l - 1] = a, b
                rec['h_k'][l - 1] = self.E if l == 1 else plan.degree_prefix(l - 1, degree, strict=False)

        if int(maf._degrees_in_host.min()) == -1:          # conditioning features feed hidden units of degree -1
            rec = np.zeros((), dtype=GROUP_DTYPE)
            hidden_rows(rec, -1)
            rows.append(rec)
        last = len(pk['groups']) - 1
        for gi, grp in enumerate(pk['groups']):
            rec = np.zeros((), dtype=GROUP_DTYPE)
            rec['out_r0'], rec['out_r1'] = grp['rows']
            rec['out_k'] = self.E if L == 1 else plan.degree_prefix(L - 1, grp['degree'], strict=True)
            rec['part_first'] = len(gparts)
            for pi, fids in enumerate(grp['ids']):
                if fids:
                    gparts.append((pi, len(ids), len(fids)))
                    ids.extend(fids)
            rec['part_count'] = len(gparts) - int(rec['part_first'])
            if gi != last:
                hidden_rows(rec, grp['degree'])
            rows.append(rec)
        self.groups_host = np.array(rows, dtype=GROUP_DTYPE) if rows else groups
        self.gparts_host = np.array(gparts, dtype=GPART_DTYPE)
        self.ids_host = np.array(ids, dtype=np.int32)
        self.max_params = max([int(r['out_r1'] - r['out_r0']) for r in rows] + [1])
        self.fixed_host = maf._fixed_indices.detach().cpu().to(torch.int32) if maf.has_fixed_indices else None
        self._dev = {}

    def _tables(self, maf, dtype, device, layouts):
        # the part records hold raw pointers to the splines' domain tensors: re-derive them when those buffers change
        dom = tuple((t.data_ptr(), t._version) for part in self.parts if part.kind == 'spline'
                    for t in (part.spec.x0, part.spec.xf, part.spec._y0, part.spec._yf))
        key = (str(device), dtype, dom)
        if key not in self._dev:
            for stale in [k for k in self._dev if k[:2] == key[:2]]:
                del self._dev[stale]
            recs = np.zeros(len(self.parts), dtype=PART_DTYPE)
            keep = []
            for i, (part, lay) in enumerate(zip(self.parts, layouts)):
                r = recs[i]
                r['kind'] = KIND[part.kind]
                spec = part.spec
                if part.kind == 'spline':
                    dom = spec.domain_tensors(dtype, device)
                    keep.append(dom)
                    r['n_bins'], r['circular'] = spec.n_bins_int, int(spec.circular)
                    r['identity_boundary_slopes'] = int(spec.identity_slopes)
                    r['learn_lower_bound'], r['learn_upper_bound'] = int(spec.learn_lower), int(spec.learn_upper)
                    r['x0'], r['xf'], r['y0'], r['yf'] = (t.data_ptr() for t in dom)
                    r['min_bin_size'], r['min_slope'] = spec.min_bin_size, spec.min_slope
                elif part.kind == 'shift':
                    tabs = spec.tables(dtype, device)
                    keep.append(tabs)
                    r['x0'], r['xf'] = tabs[0].data_ptr(), tabs[1].data_ptr()
                elif part.kind == 'moebius':
                    r['dimension'], r['unit_sphere'] = spec.dimension, int(spec.unit_sphere)
                    r['max_radius'] = float(spec.max_radius)
                cols = part.cols_on(device)
                r['cols'] = 0 if cols is None else cols.data_ptr()
                r['par_base'] = lay.base.data_ptr()
                keep.append((cols, lay.base))

            def dev(a):
                return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(device)

            self._dev[key] = dict(groups=dev(self.groups_host), parts=dev(recs), gparts=dev(self.gparts_host),
                                  ids=torch.from_numpy(self.ids_host.copy()).to(device),
                                  fixed=None if self.fixed_host is None else self.fixed_host.to(device), keep=keep,
                                  weights=None)
        return self._dev[key]

    def _padded_weights(self, maf, tb, dtype):
        """Packed effective weights with rows padded to a multiple of 16 bytes (cached by parameter version)."""
        made = maf._conditioner
        ver = made._param_versions()
        hit = tb['weights']
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        pw, pb = made.packed_weights(self.plan)
        nv = 16 // torch.empty((), dtype=dtype).element_size()
        ws = []
        for w in pw:
            pad = (-w.shape[1]) % nv
            w = w.to(dtype)
            ws.append((torch.nn.functional.pad(w, (0, pad)) if pad else w).contiguous())
        bs = [b.to(dtype).contiguous() for b in pb]
        tb['weights'] = (ver, ws, bs)
        return ws, bs

    def inverse(self, maf, y, layouts):
        """x, log_det_J of MAF.inverse(y) for a contiguous CUDA tensor (no autograd)."""
        _lib.require_cuda(y)
        y = y.contiguous()
        tb = self._tables(maf, y.dtype, y.device, layouts)
        ws, bs = self._padded_weights(maf, tb, y.dtype)
        B = y.shape[0]
        x = torch.empty_like(y)
        ld = torch.empty(B, dtype=y.dtype, device=y.device)
        if B == 0:
            return x, ld
        a = _lib.SweepArgs()
        a.dtype, a.batch, a.n_features, a.n_linear = dtype_code(y), B, self.D, self.L
        a.y, a.ldy, a.x, a.ldx, a.logdet = y.data_ptr(), y.stride(0) if B > 1 else self.D, x.data_ptr(), self.D, ld.data_ptr()
        for l in range(self.L):
            a.w[l], a.b[l] = ws[l].data_ptr(), bs[l].data_ptr()
            a.n_out[l], a.ldw[l] = ws[l].shape[0], ws[l].shape[1]
        a.groups, a.n_groups, a.max_params = tb['groups'].data_ptr(), len(self.groups_host), self.max_params
        a.parts, a.group_parts, a.ids = tb['parts'].data_ptr(), tb['gparts'].data_ptr(), tb['ids'].data_ptr()
        a.fixed_cols = None if tb['fixed'] is None else tb['fixed'].data_ptr()
        a.n_fixed = 0 if tb['fixed'] is None else len(self.fixed_host)
        a.n_embedded = self.E
        if self.embedding is not None:
            out_col, periodic = self.embedding._tables(y.device)
            a.emb_out_col, a.emb_periodic = out_col.data_ptr(), periodic.data_ptr()
            a.emb_lower, a.emb_scale = self.embedding._lower, self.embedding._scale
        ldw = [w.shape[1] for w in ws]
        a.max_group_weight_elems = max(
            [int(r['out_r1'] - r['out_r0']) * ldw[self.L - 1] +
             sum(int(r['h_b'][l - 1] - r['h_a'][l - 1]) * ldw[l - 1] for l in range(1, self.L)) for r in self.groups_host] + [0])
        with torch.cuda.device(y.device):
            check(_lib.load().tfepb_maf_inverse_sweep(ctypes.byref(a), stream_ptr(y)))
        return x, ld
