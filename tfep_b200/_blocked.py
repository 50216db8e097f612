"""Blocked inverse sweep for WIDE conditioners (BASELINE.json cfg5: D = 3000, hidden width 14998).

``MAF.inverse`` is sequential over the degrees (reference nn/flows/autoregressive.py:179-229: one full conditioner pass
per degree; made.py:366-434 gives the hidden width).  The persistent sweep kernel (tfep_b200/_sweep.py) evaluates every
unit once, but it keeps the activations of a sample tile in shared memory, which stops at ~1700 units.  Here the degrees
are cut into blocks of ``block_degrees`` consecutive degrees and the dependency structure is used twice:

* what the units of EARLIER blocks contribute to the pre-activations of block J is independent of the order inside J:
  for every linear layer one plain GEMM over the already known activations (a contiguous column prefix of the
  degree-sorted packed weights) -- these panels hold all but ~1 / n_blocks of the multiply-accumulates;
* the sequential part INSIDE block J only involves the block's own features and hidden units: a mini-network with a
  few hundred units per layer whose weights are copied out once into small padded matrices, swept degree by degree by
  the persistent kernel (tfepb_maf_inverse_sweep) with the panel results as per-sample additive terms (``extra``); the
  kernel writes the block's hidden activations straight into the full-width buffers the later panels read.

The multiply-accumulates are again nnz(masks) per sample (one forward pass) instead of n_degrees passes.  Exact fp32 /
fp64 arithmetic (the panels run on the FFMA GEMM); ``panel_precision='bf16x6'`` / ``'bf16x3'`` / ``'bf16'`` puts the
panels on the tensor cores.  Host logic here is integer work at plan time plus a loop of launches per call.
"""

import ctypes

import numpy as np
import torch

from . import _lib, _ops
from ._lib import check, dtype_code, stream_ptr
from ._sweep import GPART_DTYPE, GROUP_DTYPE, KIND, MAXL, PART_DTYPE, eligibility

#: the persistent sweep keeps (inputs + hidden units + parameters) of 32 samples in shared memory
SWEEP_MAX_UNITS = 1700


def needs_blocking(maf, pk):
    """True if the activations of a 32-sample tile of this layer do not fit the shared memory of an SM."""
    plan = pk['plan']
    return sum(len(d) for d in plan.packed_degrees[:-1]) + 64 > SWEEP_MAX_UNITS


class BlockedSweepPlan:
    def __init__(self, maf, pk, block_degrees=64):
        why = eligibility(maf, pk)
        if why is None and maf._embedding is not None:
            why = 'embeddings are not supported by the blocked sweep'
        if why is None and maf.has_fixed_indices:
            why = 'conditioning features are not supported by the blocked sweep'
        if why is not None:
            raise _lib.TfepB200Error(f'blocked inverse sweep unavailable: {why}')
        self.plan, self.parts = pk['plan'], pk['parts']
        plan = self.plan
        self.L = plan.n_layers
        self.D = len(maf._degrees_in_host)
        groups = pk['groups']
        degs = [g['degree'] for g in groups]
        assert degs == sorted(degs)
        self.widths = [len(plan.packed_degrees[l]) for l in range(self.L + 1)]
        # cumulative unit counts per degree, per hidden layer: units with degree <= d / < d as prefix lengths
        deg_arr = [plan.packed_degrees[l].numpy() for l in range(self.L + 1)]
        self.blocks = []
        gids_all = maf._groups_host
        for g0 in range(0, len(groups), block_degrees):
            g1 = min(g0 + block_degrees, len(groups))
            d_lo, d_hi = degs[g0], degs[g1 - 1]
            cols = torch.cat([gids_all[g] for g in range(g0, g1)]).tolist()          # x columns of the block, degree order
            r0, r1 = groups[g0]['rows'][0], groups[g1 - 1]['rows'][1]
            ua, ub = [0] * (self.L + 1), [0] * (self.L + 1)
            for l in range(1, self.L):
                ua[l] = int(np.searchsorted(deg_arr[l], d_lo, side='left'))
                ub[l] = int(np.searchsorted(deg_arr[l], d_hi, side='right'))
            local_col = {c: i for i, c in enumerate(cols)}
            # local group table
            recs, gparts, ids = [], [], []
            for gi in range(g0, g1):
                grp = groups[gi]
                d = grp['degree']
                rec = np.zeros((), dtype=GROUP_DTYPE)
                rec['out_r0'], rec['out_r1'] = grp['rows'][0] - r0, grp['rows'][1] - r0
                if self.L == 1:
                    rec['out_k'] = len(cols)
                else:
                    rec['out_k'] = max(0, int(np.searchsorted(deg_arr[self.L - 1], d, side='left')) - ua[self.L - 1])
                rec['part_first'] = len(gparts)
                for pi, fids in enumerate(grp['ids']):
                    if fids:
                        gparts.append((pi, len(ids), len(fids)))
                        ids.extend(fids)
                rec['part_count'] = len(gparts) - int(rec['part_first'])
                for l in range(1, self.L):
                    a = int(np.searchsorted(deg_arr[l], d, side='left')) - ua[l]
                    b = int(np.searchsorted(deg_arr[l], d, side='right')) - ua[l]
                    rec['h_a'][l - 1], rec['h_b'][l - 1] = a, b
                    rec['h_k'][l - 1] = len(cols) if l == 1 else \
                        max(0, int(np.searchsorted(deg_arr[l - 1], d, side='right')) - ua[l - 1])
                recs.append(rec)
            self.blocks.append(dict(g0=g0, g1=g1, cols=cols, rows=(r0, r1), ua=ua, ub=ub, local_col=local_col,
                                    groups=np.array(recs, dtype=GROUP_DTYPE), gparts=np.array(gparts, dtype=GPART_DTYPE),
                                    ids=np.array(ids, dtype=np.int32),
                                    max_params=max(int(r['out_r1'] - r['out_r0']) for r in recs)))
        self._dev = {}

    # -- device tables ------------------------------------------------------------------------------------------
    def _tables(self, maf, dtype, device, layouts):
        key = (str(device), dtype)
        if key in self._dev:
            return self._dev[key]

        def dev(a):
            return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(device)

        keep, per_block = [], []
        base_recs = np.zeros(len(self.parts), dtype=PART_DTYPE)
        for i, part in enumerate(self.parts):
            r = base_recs[i]
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
        part_cols = [p.x_columns() for p in self.parts]               # feature -> x column (host)
        for blk in self.blocks:
            recs = base_recs.copy()
            r0 = blk['rows'][0]
            for i, (part, lay) in enumerate(zip(self.parts, layouts)):
                # feature -> LOCAL column of the block's (batch, b) tensors; feature -> LOCAL packed output row
                lc = torch.tensor([blk['local_col'].get(int(c), 0) for c in part_cols[i].tolist()], dtype=torch.int32, device=device)
                pb = (lay.base.to(device).to(torch.int32) - r0).contiguous()
                keep.append((lc, pb))
                recs[i]['cols'], recs[i]['par_base'] = lc.data_ptr(), pb.data_ptr()
            per_block.append(dict(groups=dev(blk['groups']), parts=dev(recs), gparts=dev(blk['gparts']),
                                  ids=torch.from_numpy(blk['ids'].copy()).to(device),
                                  cols=torch.tensor(blk['cols'], dtype=torch.long, device=device)))
        self._dev[key] = dict(blocks=per_block, keep=keep, weights=None)
        return self._dev[key]

    def _block_weights(self, maf, tb, dtype):
        """Packed effective weights (full, for the panels) and the per-block mini-network matrices, padded to 16-byte
        rows; cached by parameter version."""
        made = maf._conditioner
        ver = made._param_versions()
        hit = tb['weights']
        if hit is not None and hit[0] == ver:
            return hit[1:]
        pw, pb = made.packed_weights(self.plan)
        pw = [w.to(dtype) for w in pw]
        pb = [b.to(dtype).contiguous() for b in pb]
        nv = 16 // torch.empty((), dtype=dtype).element_size()

        def padded(w):
            pad = (-w.shape[1]) % nv
            return (torch.nn.functional.pad(w, (0, pad)) if pad else w).contiguous()

        minis = []
        for blk, dv in zip(self.blocks, tb['blocks']):
            ua, ub = blk['ua'], blk['ub']
            r0, r1 = blk['rows']
            ws, bs = [], []
            for l in range(self.L):
                rows = slice(r0, r1) if l == self.L - 1 else slice(ua[l + 1], ub[l + 1])
                if l == 0:
                    w = pw[0][rows].index_select(1, dv['cols'])
                else:
                    w = pw[l][rows, ua[l]:ub[l]]
                ws.append(padded(w))
                bs.append(pb[l][rows].contiguous())
            minis.append((ws, bs))
        tb['weights'] = (ver, pw, pb, minis)
        return pw, pb, minis

    # -- the sweep ----------------------------------------------------------------------------------------------
    def inverse(self, maf, y, layouts, panel_precision='fp32'):
        """x, log_det_J of MAF.inverse(y) for a contiguous CUDA tensor (no autograd)."""
        _lib.require_cuda(y)
        y = y.contiguous()
        dtype, dev = y.dtype, y.device
        tb = self._tables(maf, dtype, dev, layouts)
        pw, pb, minis = self._block_weights(maf, tb, dtype)
        B, L = y.shape[0], self.L
        x = torch.zeros_like(y)
        ld = torch.zeros(B, dtype=dtype, device=dev)
        if B == 0:
            return x, ld
        # full-width activation buffers (zero = not known yet), read by the panels, written by the block kernels
        hidden = [None] + [torch.zeros(B, self.widths[l], dtype=dtype, device=dev) for l in range(1, L)]
        lib = _lib.load()
        for blk, dv, (ws, bs) in zip(self.blocks, tb['blocks'], minis):
            ua, ub = blk['ua'], blk['ub']
            r0, r1 = blk['rows']
            # ---- panels: what the earlier blocks contribute to this block's pre-activations ----
            extra = []
            for l in range(L):
                rows = slice(r0, r1) if l == L - 1 else slice(ua[l + 1], ub[l + 1])
                n_out = rows.stop - rows.start
                if l == 0:
                    src, w = x, pw[0][rows, :self.D]                   # unknown features are still zero in x
                else:
                    src, w = hidden[l][:, :ua[l]], pw[l][rows, :ua[l]]
                if n_out == 0 or src.shape[1] == 0 or (l == 0 and blk['g0'] == 0):
                    extra.append(None)
                    continue
                extra.append(_panel(src, w, panel_precision))
            # ---- the block's own degrees: persistent sweep over the mini-network ----
            yb = y.index_select(1, dv['cols'])
            nb = yb.shape[1]
            xb = torch.empty_like(yb)
            ldb = torch.empty(B, dtype=dtype, device=dev)
            a = _lib.SweepArgs()
            a.dtype, a.batch, a.n_features, a.n_linear = dtype_code(y), B, nb, L
            a.y, a.ldy, a.x, a.ldx, a.logdet = yb.data_ptr(), nb, xb.data_ptr(), nb, ldb.data_ptr()
            for l in range(L):
                a.w[l], a.b[l] = ws[l].data_ptr(), bs[l].data_ptr()
                a.n_out[l], a.ldw[l] = ws[l].shape[0], ws[l].shape[1]
                if extra[l] is not None:
                    a.extra[l], a.ldextra[l] = extra[l].data_ptr(), extra[l].stride(0)
                if l >= 1 and ub[l] > ua[l]:
                    a.act_out[l] = hidden[l].data_ptr() + ua[l] * hidden[l].element_size()
                    a.ldact_out[l] = hidden[l].stride(0)
            a.groups, a.n_groups, a.max_params = dv['groups'].data_ptr(), len(blk['groups']), blk['max_params']
            a.parts, a.group_parts, a.ids = dv['parts'].data_ptr(), dv['gparts'].data_ptr(), dv['ids'].data_ptr()
            a.n_fixed, a.n_embedded = 0, nb
            a.max_group_weight_elems = 0
            with torch.cuda.device(dev):
                check(lib.tfepb_maf_inverse_sweep(ctypes.byref(a), stream_ptr(y)))
            x.index_copy_(1, dv['cols'], xb)
            ld += ldb
        return x, ld


def _panel(src, w, precision):
    """src (batch, k) . w (n, k)^T without bias or activation: the FFMA GEMM, or the tensor cores."""
    if precision == 'fp32' or src.dtype != torch.float32:
        return _ops.linear_forward(src, w, None, _ops.ACT_NONE)
    n_split = {'bf16': 1, 'bf16x3': 2, 'bf16x6': 3}[precision]
    m, k = src.shape
    n = w.shape[0]
    c, _ = _ops.tc_gemm(_ops.tc_pack(src, 128, n_split=n_split), _ops.tc_pack(w, 256, n_split=n_split), m, n, k, c=True,
                        n_split=n_split)
    return c
