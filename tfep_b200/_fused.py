"""Host side of the fused tensor-core MAF layer (tfepb_maf_spline_forward_bf16).

Builds, once per MAF layer, the block schedule the kernel walks and, whenever the parameters change,
the packed bf16 weight stream:

* hidden units degree-sorted, output layer feature-major with features sorted by degree
  (tfep_b200/_pack.py), zero-padded to multiples of 16;
* every GEMM is cut into blocks of at most 24 KB (256 rows x 48 k for the output layer); a block is stored
  as the exact image of its shared-memory layout (K-major core matrices: for each group of 8 k-values, all
  rows x 16 bytes), so the bulk-copy engine moves it with one linear copy;
* biases travel inside the GEMMs: packed unit 0 and 1 of every layer input are constant ones (columns D,
  D+1 of the x operand; two extra hidden units whose own weights reproduce the one) and the weight blocks
  hold bf16(b) and bf16(b - bf16(b)) in those columns;
* each feature owns 28 consecutive output rows (25 spline parameters + 3 zero rows; 4 features = one MMA of
  N = 112); hidden-layer rows and the rows feeding softmax / softplus are pre-multiplied by log2(e), the
  hidden activations travel as log2(e) ELU(h) and the consuming layer's weights absorb the factor;
* the schedule lists only blocks that intersect the staircase of the autoregressive mask
  (``deg_out >= deg_in`` for hidden layers, ``>`` for the output layer; reference nn/masked.py:90-99):
  for a tile of output rows the reduction stops at the last input unit they may see.
"""

import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

TILE_M = 128
STAGE_BYTES = 49152
FEATS_PER_CHUNK = 4
NPAR = 25
PSTRIDE = 28
CHUNK_N = FEATS_PER_CHUNK * PSTRIDE
ACC_BUFS = 3
MAX_LAYERS, MAX_OPS = 8, 512
INV_STAGE_BYTES = 24576
INV_ACC_OUT, INV_ACC_HID, INV_A0_COL, INV_A1_COL, INV_A2_COL = 0, 32, 64, 128, 320
KB_OUT = 192
KB_HID = 112
LOG2E = 1.4426950408889634
OP_FIRST, OP_COMMIT, OP_ACC_SHIFT, OP_WAIT_A, OP_WAIT_EMPTY, OP_OWNER1, OP_HIDDEN = 1, 2, 2, 16, 32, 64, 128

OP_DTYPE = np.dtype([('w_off', '<u4'), ('idesc', '<u4'), ('n', '<u2'), ('tmem_col', '<u2'), ('a_col', '<u2'),
                     ('ksteps', 'u1'), ('flags', 'u1')])


def _idesc(n):
    """kind::f16 instruction descriptor: D = fp32 (bit 4), A = B = bf16 (bits 7, 10), K-major, N >> 3 at 17, M >> 4 at 24."""
    return (1 << 4) | (1 << 7) | (1 << 10) | ((n >> 3) << 17) | ((TILE_M >> 4) << 24)
STEP_DTYPE = np.dtype([('col', '<i4'), ('x0', '<f4'), ('period', '<f4'), ('inv_period', '<f4'), ('rw', '<f4'), ('rh', '<f4'),
                       ('y0', '<f4'), ('partner', '<i4'), ('h1_first', '<i4'), ('h1_count', '<i4'), ('h2_first', '<i4'),
                       ('h2_count', '<i4')])
FEAT_DTYPE = np.dtype([('col', '<i4'), ('x0', '<f4'), ('period', '<f4'), ('inv_period', '<f4'), ('rw', '<f4'),
                       ('rh', '<f4'), ('y0', '<f4'), ('kind', '<i4')])


def _ceil16(n):
    return (n + 15) // 16 * 16


def _spline_parts(maf):
    """The native parts of the MAF's transformer if they are all 8-bin splines the fused epilogue covers, else a reason."""
    pk = maf._pack()
    if pk is False:
        return 'transformer does not lower to native parts'
    parts = pk['parts']
    for p in parts:
        if p.kind != 'spline':
            return f'transformer part of kind {p.kind!r} (only neural splines are fused)'
        t = p.spec
        if t.n_bins_int != 8 or t.identity_slopes or t.learn_lower or t.learn_upper:
            return 'only splines with 8 bins, free boundary slopes and fixed limits are fused'
        if (t.min_bin_size, t.min_slope) != (parts[0].spec.min_bin_size, parts[0].spec.min_slope):
            return 'spline parts with different min_bin_size / min_slope'
    return parts


def eligibility(maf):
    """None if the fused kernel covers this MAF layer, else the reason it does not."""
    parts = _spline_parts(maf)
    if isinstance(parts, str):
        return parts
    from .nn.embeddings import PeriodicEmbedding
    if maf._embedding is not None and type(maf._embedding) is not PeriodicEmbedding:
        return 'only PeriodicEmbedding is fused into the operand staging'
    if maf._n_conditioner_indices > 0:
        return 'conditioner_indices are not fused'
    if len(maf._conditioner._linear_layers()) != 3:
        return 'the fused kernel is built for two hidden layers'
    D = len(maf._degrees_in_host)
    if (D * 4 * TILE_M) % 16 != 0:
        return 'row length not supported'
    return None


class FusedSplinePlan:
    def __init__(self, maf):
        why = eligibility(maf)
        if why is not None:
            raise _lib.TfepB200Error(f'fused bf16 path unavailable: {why}')
        pk = maf._pack()
        plan = pk['plan']
        parts = pk['parts']
        t = parts[0].spec                             # min_bin_size / min_slope are common to all parts
        self.D = len(maf._degrees_in_host)
        # Conditioner input columns as the kernels lay them out: the (cos, sin) pairs of the features a
        # PeriodicEmbedding lifts (reference nn/embeddings/mafembed.py:112-142) first, on even columns, then the
        # plain features, then the two constant ones.  Without embedding: input column k = x column k.
        emb = maf._embedding
        per = [] if emb is None else emb._periodic_indices.tolist()
        nonper = list(range(self.D)) if emb is None else emb._nonperiodic_indices.tolist()
        self.Din = self.D + len(per)
        self.input_of_x = {c: 2 * j for j, c in enumerate(per)}
        self.input_of_x.update({c: 2 * len(per) + i for i, c in enumerate(nonper)})
        self.lifted = set(per)
        # reference order of the embedded features (plain ones first, then the pairs) -> kernel input column
        self.input_cols = torch.tensor([self.input_of_x[c] for c in nonper] +
                                       [self.input_of_x[c] + s for c in per for s in (0, 1)], dtype=torch.long)
        self.K1 = _ceil16(self.Din + 2)               # + two constant-one columns carrying the bias
        if emb is None:
            self.input_map_host, self.emb_lower, self.emb_scale = None, 0.0, 0.0
        else:
            imap = np.full(self.K1, 4 << 16, dtype=np.int32)
            for c in range(self.D):
                k = self.input_of_x[c]
                if c in self.lifted:
                    imap[k], imap[k + 1] = c | (1 << 16), c | (2 << 16)
                else:
                    imap[k] = c
            imap[self.Din] = imap[self.Din + 1] = 3 << 16
            self.input_map_host, self.emb_lower, self.emb_scale = imap, emb._lower, emb._scale
        deg_h1, deg_h2 = plan.packed_degrees[1], plan.packed_degrees[2]
        H1, H2 = len(deg_h1), len(deg_h2)
        if H1 != H2:
            raise _lib.TfepB200Error('fused bf16 path needs equal hidden widths')
        self.H = H1
        self.HP = _ceil16(H1 + 2)                     # packed units 0, 1 are the constant ones
        if self.HP > 336 or self.K1 > 352:
            raise _lib.TfepB200Error('layer widths exceed the tensor-memory plan of the fused kernel (at most 334 hidden units '
                                     'and 350 conditioner inputs: e.g. MAF(..., hidden_layers=[334, 334]))')
        self.perm1, self.perm2 = plan.perms[1], plan.perms[2]

        # Hidden layers are computed and handed over in two column halves (rows of the GEMM = columns of the
        # accumulator).  Layer 1 splits at s1; layer 2 splits at the largest s2 whose rows only see layer-1 units
        # below s1, so that GEMM2 of the first half can run while the ELU of the second half of h1 is in progress.
        def kmax2(rows_end):
            real = deg_h2[:max(min(rows_end - 2, self.H), 0)]
            return _ceil16(2 + int((deg_h1 <= int(real.max())).sum())) if len(real) else 16

        s1 = 16 * math.ceil(self.HP / 32)
        s2 = s1
        while s2 > 16 and kmax2(s2) > s1:
            s2 -= 16
        self.halves = 2 if (self.HP >= 64 and kmax2(s2) <= s1 and (self.HP - s2) * KB_HID * 2 <= STAGE_BYTES) else 1
        if self.halves == 2:
            self.hidden_chunks1 = [(0, s1), (s1, self.HP)]
            self.hidden_chunks2 = [(0, s2), (s2, self.HP)]
        else:
            self.hidden_chunks1 = self.hidden_chunks2 = [(0, self.HP)]
        self.hidden_split = (self.hidden_chunks1[0][1], self.hidden_chunks2[0][1])

        # features of all spline parts (a MixedTransformer contributes one part per child), sorted by degree ->
        # chunks of 4 slots; a feature is (x column, degree, spline module, local index, 25 reference output rows)
        deg_in = maf._degrees_in_host
        feats_all = []
        for part in parts:
            cols_p = part.x_columns().tolist()
            ref_cols_p = part.ref_columns()                 # (F, 25) rows of the reference output layer
            dom = [b.detach().float().cpu() for b in (part.spec.x0, part.spec.xf, part.spec._y0, part.spec._yf)]
            for f in range(part.n_features):
                feats_all.append(dict(col=cols_p[f], deg=int(deg_in[cols_p[f]]), rows=ref_cols_p[f], circular=part.spec.circular,
                                      x0=float(dom[0][f]), xf=float(dom[1][f]), y0=float(dom[2][f]), yf=float(dom[3][f])))
        feats_all.sort(key=lambda d: (d['deg'], d['col']))
        # conditioning features (degree -1) enter the conditioner and pass through: no slot, no output rows
        if sorted(d['col'] for d in feats_all) != [c for c in range(self.D) if int(deg_in[c]) >= 0]:
            raise _lib.TfepB200Error('fused bf16 path: the spline parts must cover every mapped feature exactly once')
        self._feats_all = feats_all
        self.mixed = any(not d['circular'] for d in feats_all)
        order = list(range(len(feats_all)))
        cols = [d['col'] for d in feats_all]            # indexed by sorted position
        self._order = order
        self._inv, self._inv_why = None, None
        self.n_chunks = math.ceil(len(order) / FEATS_PER_CHUNK)
        feats = np.zeros(self.n_chunks * FEATS_PER_CHUNK, dtype=FEAT_DTYPE)
        feats['col'] = -1
        w3_rows = torch.full((self.n_chunks * CHUNK_N,), -1, dtype=torch.long)     # -1 -> zero row
        chunk_maxdeg = []
        scale = torch.ones(len(w3_rows))
        for slot, d in enumerate(feats_all):
            c, j = divmod(slot, FEATS_PER_CHUNK)
            L = d['xf'] - d['x0']
            mi = 8 * t.min_bin_size
            feats[slot] = (d['col'], d['x0'], L, 1.0 / L, np.float32(L) - np.float32(mi),
                           np.float32(d['yf'] - d['y0']) - np.float32(mi), d['y0'], 0 if d['circular'] else 1)
            r0 = c * CHUNK_N + j * PSTRIDE
            w3_rows[r0:r0 + NPAR] = d['rows']
            # widths, heights, slopes live in the log2 domain; the shift of a circular spline does not
            scale[r0:r0 + (24 if d['circular'] else 25)] = LOG2E
        for c in range(self.n_chunks):
            fs = order[c * FEATS_PER_CHUNK:(c + 1) * FEATS_PER_CHUNK]
            chunk_maxdeg.append(max(int(deg_in[cols[f]]) for f in fs))
        self.feats_host = feats
        self.w3_rows = w3_rows
        self.w3_scale = scale
        assert self.n_chunks * CHUNK_N == len(w3_rows)

        # the first output chunk must live on the first half of h2 (see GEMM3 below); narrow layers are not split
        if self.halves == 2 and (_ceil16(2 + int((deg_h2 < chunk_maxdeg[0]).sum())) > self.hidden_split[1] or
                                 CHUNK_N > self.hidden_split[1] or self.n_chunks < 2):
            self.halves = 1
            self.hidden_chunks1 = self.hidden_chunks2 = [(0, self.HP)]
            self.hidden_split = (self.HP, self.HP)
        if any((b - a) * self.K1 * 2 > STAGE_BYTES for a, b in self.hidden_chunks1):
            raise _lib.TfepB200Error('fused bf16 path: first-layer weight block exceeds a ring stage')

        # x sits at the end of the A operand region (176 tensor-memory columns, two k-values each), clear of the first
        # half of h1, which is then written while the second half of GEMM1 still reads x
        self.x_col = (176 - self.K1 // 2) // 8 * 8

        # ---- schedule ----
        ops, gather = [], []           # gather: per op, index tensor into the concatenated padded matrices
        off1, off2 = self.HP * self.K1, self.HP * self.K1 + self.HP * self.HP
        w_off = 0

        def add(n, tmem_col, kb, kmax_block, a_slab0, flags, base, ld, r0):
            nonlocal w_off
            ksteps = kmax_block // 16
            rows = torch.arange(r0, r0 + n)
            ks = torch.arange(kb, kb + kmax_block)
            idx = (base + rows[:, None] * ld + ks[None, :]).reshape(n, 2 * ksteps, 8).permute(1, 0, 2).reshape(-1)
            gather.append(idx)
            nbytes = n * kmax_block * 2
            assert nbytes <= STAGE_BYTES and nbytes % 16 == 0
            ops.append((w_off, _idesc(n), n, tmem_col, a_slab0 * 4, ksteps, flags))
            w_off += nbytes

        # Accumulator groups alternate between the two MMA issuer warps (OP_OWNER1); every group ends with a
        # commit.  Hidden layers: one group per column half (its index travels in the accumulator bits); every
        # OP_WAIT_A consumes the next hand-over of the epilogue, in the order x, h1 half 0, h1 half 1, h2 half 0,
        # h2 half 1.
        # GEMM1: full K1 per half (the first layer is tiny; no staircase); both halves need only x
        for i, (a, b) in enumerate(self.hidden_chunks1):
            fl = OP_FIRST | OP_COMMIT | OP_HIDDEN | (i << OP_ACC_SHIFT) | (OP_WAIT_A if i == 0 else 0) | (OP_OWNER1 if i & 1 else 0)
            add(b - a, a, 0, self.K1, self.x_col // 4, fl, 0, self.K1, a)
        # GEMM2: rows see layer-1 units of degree <= their own; half i waits for half i of h1
        for i, (a, b) in enumerate(self.hidden_chunks2):
            kmax = kmax2(b)
            if self.halves == 2 and i == 0:
                assert kmax <= self.hidden_split[0]
            blocks = list(range(0, kmax, KB_HID))
            for bi, kb in enumerate(blocks):
                fl = OP_HIDDEN | (i << OP_ACC_SHIFT) | (OP_FIRST | OP_WAIT_A if bi == 0 else 0) | (OP_OWNER1 if i & 1 else 0)
                if bi == len(blocks) - 1:
                    fl |= OP_COMMIT
                add(b - a, a, kb, min(KB_HID, kmax - kb), kb // 8, fl, off1, self.HP, a)
        # GEMM3: a chunk of features sees layer-2 units of degree < its largest degree.  The first chunk only needs
        # the first half of h2 (its accumulator lies inside the columns that half has released); the second chunk
        # consumes the hand-over of the second half, after which everything is available.
        for c in range(self.n_chunks):
            kmax = _ceil16(2 + int((deg_h2 < chunk_maxdeg[c]).sum()))
            blocks = list(range(0, kmax, KB_OUT))
            acc = c % ACC_BUFS
            wait_a = c == 0 or (c == 1 and self.halves == 2)
            if self.halves == 2 and c == 0 and (kmax > self.hidden_split[1] or CHUNK_N > self.hidden_split[1] or self.n_chunks < 2):
                raise _lib.TfepB200Error('fused bf16 path: first output chunk does not fit the first half of h2')
            for bi, kb in enumerate(blocks):
                fl = (acc << OP_ACC_SHIFT) | (OP_OWNER1 if c & 1 else 0)
                if bi == 0:
                    fl |= OP_FIRST | OP_WAIT_EMPTY | (OP_WAIT_A if wait_a else 0)
                if bi == len(blocks) - 1:
                    fl |= OP_COMMIT
                add(CHUNK_N, acc * CHUNK_N, kb, min(KB_OUT, kmax - kb), kb // 8, fl, off2, self.HP, c * CHUNK_N)
        self.ops_host = np.array(ops, dtype=OP_DTYPE)
        self.gather_host = torch.cat(gather)
        self.weight_bytes = w_off
        self.mma_columns = int(sum(int(o[2]) * int(o[5]) for o in ops))          # sum of N x ksteps over all blocks
        self.min_bin, self.min_slope = float(t.min_bin_size), float(t.min_slope)
        self.slope_offset = float(math.log(math.exp(1.0 - t.min_slope) - 1.0))
        self._dev = {}
        self._cache = None

    # ---------------------------------------------------------------------------------------------
    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            feats = torch.from_numpy(self.feats_host.view(np.uint8).copy()).to(device)
            imap = None if self.input_map_host is None else torch.from_numpy(self.input_map_host).to(device)
            self._dev[key] = dict(feats=feats, gather=self.gather_host.to(device), input_map=imap, input_cols=self.input_cols.to(device),
                                  w3_rows=self.w3_rows.to(device), w3_scale=self.w3_scale.to(device),
                                  perm1=self.perm1.to(device),
                                  perm2=self.perm2.to(device), err=torch.zeros(1, dtype=torch.int32, device=device))
        return self._dev[key]

    def pack(self, maf):
        """Packed bf16 weight stream + fp32 biases from the current parameters (cached by parameter version)."""
        made = maf._conditioner
        key = made._param_versions()
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        with torch.no_grad():
            (w1, b1), (w2, b2), (w3, b3) = made.effective_weights()
            dev = w1.device
            tb = self._tables(dev)
            H, HP, K1, D = self.H, self.HP, self.K1, self.Din      # D: conditioner inputs (lifted pairs included)

            def hi_lo(b):
                hi = b.to(torch.bfloat16).float()
                return hi, b - hi

            # The accumulators of the hidden layers hold t = log2(e) (W a + b) and the epilogue stores
            # a' = log2(e) ELU(t / log2(e)) (one ex2 + one fma per unit), so: layer 1 is scaled by log2(e);
            # layer 2 sees a' and is scaled by log2(e) itself -> weights unchanged, bias scaled; the output rows
            # see a' -> divide by log2(e), then the log2-domain rows are multiplied by it again.
            W1p = torch.zeros(HP, K1, device=dev)
            W1p[2:2 + H, tb['input_cols']] = w1.index_select(0, tb['perm1']) * LOG2E
            W1p[2:2 + H, D], W1p[2:2 + H, D + 1] = hi_lo(b1.index_select(0, tb['perm1']) * LOG2E)
            W1p[0, D] = W1p[1, D] = 1.0                   # hidden units 0, 1: t = 1 > 0 -> a' = 1 (constant ones)
            W2p = torch.zeros(HP, HP, device=dev)
            W2p[2:2 + H, 2:2 + H] = w2.index_select(0, tb['perm2']).index_select(1, tb['perm1'])
            W2p[2:2 + H, 0], W2p[2:2 + H, 1] = hi_lo(b2.index_select(0, tb['perm2']) * LOG2E)
            W2p[0, 0] = W2p[1, 0] = 1.0
            rows, scale = tb['w3_rows'], tb['w3_scale']
            safe = torch.where(rows < 0, torch.full_like(rows, w3.shape[0]), rows)
            w3e = torch.cat([w3, torch.zeros(1, H, device=dev)], dim=0)
            b3e = torch.cat([b3, torch.zeros(1, device=dev)])
            W3p = torch.zeros(len(rows), HP, device=dev)
            W3p[:, 2:2 + H] = w3e.index_select(0, safe).index_select(1, tb['perm2']) * (scale / LOG2E)[:, None]
            W3p[:, 0], W3p[:, 1] = hi_lo(b3e.index_select(0, safe) * scale)
            # the trailing zero is what padding rows of the inverse blocks point at
            src = torch.cat([W1p.flatten(), W2p.flatten(), W3p.flatten(), torch.zeros(1, device=dev)]).to(torch.bfloat16)
            packed = src.index_select(0, tb['gather']).contiguous()
        self._cache = (key, packed, src)
        return self._cache[1]

    # ---------------------------------------------------------------------------------------------
    # inverse direction (tfepb_maf_spline_inverse_bf16): one block per product of the degree sweep
    # ---------------------------------------------------------------------------------------------
    def inverse_eligibility(self, maf):
        """None if the tensor-core inverse sweep covers this layer, else the reason it does not."""
        if self._inv is None:
            self._build_inverse(maf)
        return self._inv_why if self._inv is False else None

    def _build_inverse(self, maf):
        """Schedule, step table and gather indices of the degree-ordered sweep (host integer work, once)."""
        pk = maf._pack()
        plan = pk['plan']
        deg_in = maf._degrees_in_host
        deg_h1, deg_h2 = plan.packed_degrees[1], plan.packed_degrees[2]
        cols = [d['col'] for d in self._feats_all]
        order = self._order
        why = None
        degs = [d['deg'] for d in self._feats_all]
        if why is None and any(b <= a for a, b in zip(degs, degs[1:])):
            why = 'more than one feature per degree'
        if self.K1 > 128 or self.HP > 352:
            why = 'layer widths exceed the tensor-memory plan of the inverse kernel'
        ops, steps, gather = [], [], []
        off1, off2 = self.HP * self.K1, self.HP * self.K1 + self.HP * self.HP
        zero_idx = off2 + len(self.w3_rows) * self.HP
        w_off = 0

        def add(rows, n, kmax, a_col, tmem_col, base, ld):
            nonlocal w_off
            ksteps = kmax // 16
            r = torch.full((n,), -1, dtype=torch.long)
            r[:len(rows)] = torch.as_tensor(rows, dtype=torch.long)
            ks = torch.arange(kmax)
            idx = base + r[:, None] * ld + ks[None, :]
            idx = torch.where(r[:, None] < 0, torch.full_like(idx, zero_idx), idx)
            gather.append(idx.reshape(n, 2 * ksteps, 8).permute(1, 0, 2).reshape(-1))
            nbytes = n * kmax * 2
            assert nbytes <= INV_STAGE_BYTES and nbytes % 16 == 0
            ops.append((w_off, _idesc(n), n, tmem_col, a_col, ksteps, 0))
            w_off += nbytes

        # Conditioning features (degree -1) are known from the start: the kernel stages them into the x operand from
        # `init_map`, and leading steps without a feature (col = -1) bring in the hidden units that depend on them
        # alone, in chunks of at most 15 units (first all of layer 1, then layer 2, which sees every such unit).
        cond = [c for c in range(self.D) if int(deg_in[c]) < 0]
        init_map = None
        if cond and why is None:
            init_map = np.full(self.K1, 4 << 16, dtype=np.int32)
            for c in cond:
                k = self.input_of_x[c]
                if c in self.lifted:
                    init_map[k], init_map[k + 1] = c | (1 << 16), c | (2 << 16)
                else:
                    init_map[k] = c
            init_map[self.Din] = init_map[self.Din + 1] = 3 << 16
            for layer, (deg_h, a_col, base, ld) in enumerate(((deg_h1, INV_A0_COL, 0, self.K1), (deg_h2, INV_A1_COL, off1, self.HP))):
                units = (deg_h == -1).nonzero().flatten()
                kmax = self.K1 if layer == 0 else _ceil16(2 + int((deg_h1 <= -1).sum()))
                for i in range(0, len(units), 15):
                    first, n = 2 + int(units[i]), len(units[i:i + 15])
                    add(list(range(first, first + n)), 16, kmax, a_col, INV_ACC_HID, base, ld)
                    h = (first, n, 0, 0) if layer == 0 else (0, 0, first, n)
                    steps.append((-1, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0) + h)
        for si, f in enumerate(order if why is None else []):
            d = degs[si]
            c, j = divmod(si, FEATS_PER_CHUNK)
            r0 = c * CHUNK_N + j * PSTRIDE
            kmax = _ceil16(2 + int((deg_h2 < d).sum()))
            add(list(range(r0, r0 + NPAR)), 32, kmax, INV_A2_COL, INV_ACC_OUT, off2, self.HP)
            u1 = (deg_h1 == d).nonzero().flatten()
            u2 = (deg_h2 == d).nonzero().flatten()
            last = si == len(order) - 1
            h1 = (0, 0) if (last or len(u1) == 0) else (2 + int(u1[0]), len(u1))
            h2 = (0, 0) if (last or len(u2) == 0) else (2 + int(u2[0]), len(u2))
            if h1[1] > 15 or h2[1] > 15:
                why = 'more than 15 hidden units of one degree'
                break
            if h1[1]:
                assert torch.equal(u1, torch.arange(int(u1[0]), int(u1[0]) + len(u1)))
                add(list(range(h1[0], h1[0] + h1[1])), 16, self.K1, INV_A0_COL, INV_ACC_HID, 0, self.K1)
            if h2[1]:
                assert torch.equal(u2, torch.arange(int(u2[0]), int(u2[0]) + len(u2)))
                kmax2 = _ceil16(2 + int((deg_h1 <= d).sum()))
                add(list(range(h2[0], h2[0] + h2[1])), 16, kmax2, INV_A1_COL, INV_ACC_HID, off1, self.HP)
            col = cols[f]
            ic = self.input_of_x[col]                      # conditioner input column of the feature
            if col in self.lifted:
                partner = 32 | (ic << 8)                   # (cos, sin) fill the pair column on their own
            else:
                x_of_input = {k: c for c, k in self.input_of_x.items() if c not in self.lifted}
                pi = ic ^ 1
                pc = x_of_input.get(pi, 0)
                partner = 2 if pi == self.Din else (1 if (pi in x_of_input and int(deg_in[pc]) < d) else 0)
                partner |= (ic << 8) | (pc << 16)
            partner |= 0 if self._feats_all[si]['circular'] else 16
            ft = self.feats_host[si]
            steps.append((col, ft['x0'], ft['period'], ft['inv_period'], ft['rw'], ft['rh'], ft['y0'], partner,
                          h1[0], h1[1], h2[0], h2[1]))
        if why is not None:
            self._inv, self._inv_why = False, why
            return
        self._inv = dict(ops=np.array(ops, dtype=OP_DTYPE), steps=np.array(steps, dtype=STEP_DTYPE),
                         gather=torch.cat(gather), init_map=init_map, dev={}, cache=None)

    def inverse_tables(self, maf, device):
        """(ops, steps, packed weights) on the device for the current parameters."""
        if self._inv is None:
            self._build_inverse(maf)
        if self._inv is False:
            raise _lib.TfepB200Error(f'fused bf16 inverse unavailable: {self._inv_why}')
        inv = self._inv
        key = str(device)
        if key not in inv['dev']:
            inv['dev'][key] = dict(ops=torch.from_numpy(inv['ops'].view(np.uint8).reshape(-1).copy()).to(device),
                                   steps=torch.from_numpy(inv['steps'].view(np.uint8).reshape(-1).copy()).to(device),
                                   gather=inv['gather'].to(device),
                                   init_map=None if inv['init_map'] is None else torch.from_numpy(inv['init_map']).to(device))
        tb = inv['dev'][key]
        self.pack(maf)                                   # refreshes the padded source matrices if parameters changed
        ver, _, src = self._cache
        if inv['cache'] is None or inv['cache'][0] != (ver, key):
            inv['cache'] = ((ver, key), src.index_select(0, tb['gather']).contiguous())
        return tb['ops'], tb['steps'], inv['cache'][1], tb['init_map']

    def forward(self, maf, x, debug_params=None):
        """y, log_det_J = fused layer on a contiguous fp32 CUDA tensor (no autograd)."""
        return run_chain([(self, maf)], x, debug_params=debug_params)


_EPOCH = [0]


def _chain_key(pl):
    return (pl.D, pl.Din, pl.K1, pl.HP, pl.halves, pl.x_col)


def chain_compatible(plans):
    """One launch serves a chain of layers only if they share the widths (the split of the hidden layers is per layer)."""
    return all(_chain_key(pl) == _chain_key(plans[0]) for pl in plans)


LAYERS_PER_LAUNCH = 4          # the feature tables of all layers of a launch share the 227 KB of shared memory


def run_chain(plans_mafs, x, debug_params=None):
    """A chain of fused MAF layers, LAYERS_PER_LAUNCH per launch: y, sum of log_det_J (reference sequential.py:50-68)."""
    if len(plans_mafs) > LAYERS_PER_LAUNCH:
        y, ld = x, None
        for i in range(0, len(plans_mafs), LAYERS_PER_LAUNCH):
            y, l = run_chain(plans_mafs[i:i + LAYERS_PER_LAUNCH], y)
            ld = l if ld is None else ld + l
        return y, ld
    _lib.require_cuda(x)
    if x.dtype != torch.float32:
        raise _lib.TfepB200Error('the fused bf16 path takes float32 inputs')
    n_layers = len(plans_mafs)
    first = plans_mafs[0][0]
    if n_layers > MAX_LAYERS or sum(len(pl.ops_host) for pl, _ in plans_mafs) > MAX_OPS:
        raise _lib.TfepB200Error('chain too long for one fused launch')
    if not chain_compatible([pl for pl, _ in plans_mafs]):
        raise _lib.TfepB200Error('fused chain needs layers of identical widths')
    x = x.contiguous()
    B = x.shape[0]
    y = torch.empty_like(x)
    ld = torch.empty(B, dtype=torch.float32, device=x.device)
    if B == 0:
        return y, ld
    layers = (_lib.FusedLayer * n_layers)()
    keep = []
    for i, (pl, maf) in enumerate(plans_mafs):
        packed = pl.pack(maf)
        tb = pl._tables(x.device)
        keep.append(packed)
        layers[i] = _lib.FusedLayer(ops=pl.ops_host.ctypes.data, n_ops=len(pl.ops_host), n_chunks=pl.n_chunks,
                                    weights=packed.data_ptr(), feats=tb['feats'].data_ptr(), min_bin_size=pl.min_bin,
                                    min_slope=pl.min_slope, slope_offset=pl.slope_offset, reserved=0,
                                    input_map=None if tb['input_map'] is None else tb['input_map'].data_ptr(),
                                    emb_lower=pl.emb_lower, emb_scale=pl.emb_scale,
                                    hidden_split=(ctypes.c_int32 * 2)(*pl.hidden_split))
    tb = first._tables(x.device)
    flags = None
    if n_layers > 1:
        # per-tile publication flags; a fresh epoch per launch instead of a memset (launches using the same
        # workspace are serialised on the stream)
        need = (n_layers - 1) * ((B + TILE_M - 1) // TILE_M)
        key = ('flags', torch.cuda.current_stream(x.device).cuda_stream)
        flags = tb.get(key)
        if flags is None or flags.numel() < need:
            flags = tb[key] = torch.zeros(max(need, 4096), dtype=torch.int32, device=x.device)
    _EPOCH[0] = (_EPOCH[0] % 0x7fffffff) + 1
    if flags is not None and torch.cuda.is_current_stream_capturing():
        flags.zero_()           # a replayed CUDA graph re-uses the captured epoch: reset the flags inside the graph
    args = _lib.FusedArgs(x=x.data_ptr(), y=y.data_ptr(), logdet=ld.data_ptr(), batch=B, n_features=first.D,
                          k1=first.K1, hidden_padded=first.HP, n_layers=n_layers, hidden_halves=first.halves,
                          hidden_split=(ctypes.c_int32 * 2)(*first.hidden_split), layers=layers,
                          tile_flags=None if flags is None else flags.data_ptr(), epoch=_EPOCH[0],
                          debug_mode=int(os.environ.get('TFEPB_FUSED_DEBUG_MODE', '0')),
                          mixed_splines=int(any(pl.mixed for pl, _ in plans_mafs)), n_inputs=first.Din,
                          x_operand_column=first.x_col, reserved3=0,
                          error_flag=tb['err'].data_ptr(),
                          debug_params=None if debug_params is None else debug_params.data_ptr())
    with torch.cuda.device(x.device):
        check(_lib.load().tfepb_maf_spline_forward_bf16(ctypes.byref(args), stream_ptr(x)))
    return y, ld


def run_inverse_chain(plans_mafs, y):
    """One launch for the inverse of a chain of fused MAF layers, given in the order they are inverted:
    x, sum of log_det_J (reference sequential.py:50-68 with inverse=True)."""
    if len(plans_mafs) > MAX_LAYERS:
        x, ld = y, None
        for i in range(0, len(plans_mafs), MAX_LAYERS):
            x, l = run_inverse_chain(plans_mafs[i:i + MAX_LAYERS], x)
            ld = l if ld is None else ld + l
        return x, ld
    _lib.require_cuda(y)
    if y.dtype != torch.float32:
        raise _lib.TfepB200Error('the fused bf16 path takes float32 inputs')
    n_layers = len(plans_mafs)
    first = plans_mafs[0][0]
    if any((pl.D, pl.Din, pl.K1, pl.HP) != (first.D, first.Din, first.K1, first.HP) for pl, _ in plans_mafs):
        raise _lib.TfepB200Error('fused chain needs layers of identical widths')
    y = y.contiguous()
    B = y.shape[0]
    x = torch.empty_like(y)
    ld = torch.empty(B, dtype=torch.float32, device=y.device)
    if B == 0:
        return x, ld
    layers = (_lib.FusedInvLayer * n_layers)()
    keep = []
    for i, (pl, maf) in enumerate(plans_mafs):
        ops, steps, packed, init_map = pl.inverse_tables(maf, y.device)
        keep.append((ops, steps, packed, init_map))
        layers[i] = _lib.FusedInvLayer(ops=ops.data_ptr(), steps=steps.data_ptr(), n_ops=len(pl._inv['ops']),
                                       n_steps=len(pl._inv['steps']), weights=packed.data_ptr(), min_bin_size=pl.min_bin,
                                       min_slope=pl.min_slope, slope_offset=pl.slope_offset, reserved=0,
                                       emb_lower=pl.emb_lower, emb_scale=pl.emb_scale,
                                       init_map=None if init_map is None else init_map.data_ptr())
    tb = first._tables(y.device)
    flags = None
    if n_layers > 1:
        need = (n_layers - 1) * ((B + TILE_M - 1) // TILE_M)
        key = ('flags', torch.cuda.current_stream(y.device).cuda_stream)
        flags = tb.get(key)
        if flags is None or flags.numel() < need:
            flags = tb[key] = torch.zeros(max(need, 4096), dtype=torch.int32, device=y.device)
    _EPOCH[0] = (_EPOCH[0] % 0x7fffffff) + 1
    if flags is not None and torch.cuda.is_current_stream_capturing():
        flags.zero_()
    args = _lib.FusedInvArgs(y=y.data_ptr(), x=x.data_ptr(), logdet=ld.data_ptr(), batch=B, n_features=first.D,
                             k1=first.K1, hidden_padded=first.HP, n_layers=n_layers,
                             mixed_splines=int(any(pl.mixed for pl, _ in plans_mafs)), layers=layers,
                             tile_flags=None if flags is None else flags.data_ptr(), epoch=_EPOCH[0], n_inputs=first.Din,
                             error_flag=tb['err'].data_ptr())
    with torch.cuda.device(y.device):
        check(_lib.load().tfepb_maf_spline_inverse_bf16(ctypes.byref(args), stream_ptr(y)))
    return x, ld
