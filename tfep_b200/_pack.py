"""Degree-sorted packing of a MADE conditioner (host logic, integer work at construction time).

Hidden units of a MADE can be relabelled freely (permute the rows of layer l and the columns of
layer l+1 identically).  Sorting every layer's units by autoregressive degree turns each mask
(``deg_out >= deg_in`` for hidden layers, ``>`` for the output layer; reference nn/masked.py:90-99,
nn/conditioners/made.py:308-309) into a block lower-triangular staircase, so that

  * for a tile of consecutive output units only a bounded range [begin, end) of the reduction
    dimension holds non-zero weights -- the rest is never read or multiplied (``k_ranges`` forward,
    ``n_ranges`` for the transposed product of the backward pass);
  * all units of one degree are contiguous, which is what the degree-ordered inverse sweep needs.

In the reference's native order hidden degrees are a round-robin tile of the input degrees
(made.py:413-419) and no tile is empty.  The permutation is invisible outside: parameters and
``state_dict`` stay in the reference's order, packed copies are derived tensors.
"""

import torch

from ._lib import GEMM_TILE_N


def _stable_argsort(deg):
    return torch.sort(deg, stable=True).indices


def _ranges(packed_mask, tile, axis):
    """Bounding [begin, end) of the non-zeros of ``packed_mask`` for every tile of `tile` rows
    (axis=0: per row tile, range over columns) or columns (axis=1: range over rows)."""
    m = packed_mask if axis == 0 else packed_mask.t()
    n_tiles = (m.shape[0] + tile - 1) // tile
    out = torch.zeros(n_tiles, 2, dtype=torch.int32)
    for t in range(n_tiles):
        nz = m[t * tile:(t + 1) * tile].any(dim=0).nonzero().flatten()
        if len(nz) > 0:
            out[t, 0], out[t, 1] = int(nz[0]), int(nz[-1]) + 1
    return out


class MadePlan:
    """Permutations, reduction ranges and per-degree row groups of a packed MADE.

    Parameters
    ----------
    degree_chain : list of LongTensor
        Degrees of the input, every hidden layer and the output, in the reference's order.
    out_order : LongTensor, optional
        Order in which the OUTPUT units are packed (indices into the reference's output order).
        Default: the reference's order (what ``MADE.forward`` must return).
    """

    def __init__(self, degree_chain, out_order=None):
        chain = [torch.as_tensor(d).long().cpu() for d in degree_chain]
        self.n_layers = len(chain) - 1
        self.perms = [torch.arange(len(chain[0]))]
        for l in range(1, self.n_layers):
            self.perms.append(_stable_argsort(chain[l]))
        self.perms.append(torch.arange(len(chain[-1])) if out_order is None else torch.as_tensor(out_order).long().cpu())
        # out_order may hold -1 entries: all-zero padding rows of the packed output layer (tfep_b200/_txfused.py); their
        # "degree" is below every input degree, so the masks, ranges and counts below see them as empty rows
        last = self.perms[-1]
        self.has_padding = bool((last < 0).any())
        self.packed_degrees = [chain[l][self.perms[l]] for l in range(self.n_layers)]
        self.packed_degrees.append(torch.where(last >= 0, chain[-1][last.clamp(min=0)], torch.full_like(last, -(1 << 30))))
        self.k_ranges, self.n_ranges, self.nnz = [], [], []
        for l in range(self.n_layers):
            d_in, d_out = self.packed_degrees[l], self.packed_degrees[l + 1]
            strict = l == self.n_layers - 1
            mask = (d_out[:, None] > d_in[None, :]) if strict else (d_out[:, None] >= d_in[None, :])
            self.k_ranges.append(_ranges(mask, GEMM_TILE_N, axis=0))
            self.n_ranges.append(_ranges(mask, GEMM_TILE_N, axis=1))
            self.nnz.append(int(mask.sum()))
        self._device_cache = {}

    @property
    def masked_macs(self):
        """Multiply-accumulates per sample that the masks leave (the algorithmic work of one pass)."""
        return sum(self.nnz)

    def tables(self, device):
        """(k_ranges, n_ranges) as lists of int32 device tensors."""
        key = str(device)
        if key not in self._device_cache:
            perms = [p.to(device) for p in self.perms]
            if self.has_padding:
                n_out = len(self.perms[-1]) - int((self.perms[-1] < 0).sum())      # = index of the appended zero row
                perms[-1] = torch.where(perms[-1] >= 0, perms[-1], torch.full_like(perms[-1], n_out))
            self._device_cache[key] = ([r.to(device) for r in self.k_ranges], [r.to(device) for r in self.n_ranges], perms)
        return self._device_cache[key]

    @staticmethod
    def tc_ranges_of_mask(mask, device):
        """(forward k-block ranges, backward-input k-block ranges, weight-gradient row ranges) of one masked layer for the
        tensor-core GEMM; see :meth:`tc_ranges`.  ``mask``: (outputs, inputs) bool."""
        out = []
        for m in (mask, mask.t()):
            r = _ranges(m, 256, axis=0)                    # per 256-row tile: bounding column range
            kb = torch.stack([r[:, 0] // 64, (r[:, 1] + 63) // 64], dim=1).to(torch.int32)
            out.append(kb.contiguous().to(device))
        # weight gradient: the (128-output-row, 256-input-column) tiles the mask leaves non-zero, as a list
        from ._ops import WgTiles
        n_out, n_in = mask.shape
        tiles = [(tm, tn) for tm in range((n_out + 127) // 128) for tn in range((n_in + 255) // 256)
                 if bool(mask[tm * 128:(tm + 1) * 128, tn * 256:(tn + 1) * 256].any())]
        out.append(WgTiles(torch.tensor(tiles, dtype=torch.int32).reshape(-1, 2).contiguous().to(device)))
        return tuple(out)

    def tc_ranges(self, device):
        """Per layer, for the tensor-core GEMM (tiles of 256 output columns, k-blocks of 64): the non-zero k-block
        range [first, end) of every tile, forward (tiles over the layer's outputs, k over its inputs) and backward
        input (tiles over the inputs, k over the outputs), and for the weight gradient the row range of every tile of 256
        input columns (as ``_ops.WgTiles``: the list of non-zero tiles).  Lists of int32 (tiles, 2) device tensors."""
        key = ('tc', str(device))
        if key not in self._device_cache:
            fwd, bwd, roww = [], [], []
            for l in range(self.n_layers):
                d_in, d_out = self.packed_degrees[l], self.packed_degrees[l + 1]
                strict = l == self.n_layers - 1
                mask = (d_out[:, None] > d_in[None, :]) if strict else (d_out[:, None] >= d_in[None, :])
                f, b, r = self.tc_ranges_of_mask(mask, device)
                fwd.append(f)
                bwd.append(b)
                roww.append(r)
            self._device_cache[key] = (fwd, bwd, roww)
        return self._device_cache[key]

    def perm_tables(self, device):
        """int32 device copies of the permutations for tfepb_wn_pack: rows of layer l are ``[l + 1]`` (-1 = zero row),
        columns ``[l]`` (None for the identity of the input layer)."""
        key = ('perm32', str(device))
        if key not in self._device_cache:
            t = [p.to(device=device, dtype=torch.int32) for p in self.perms]
            t[0] = None
            self._device_cache[key] = t
        return self._device_cache[key]

    def pack(self, weights, biases):
        """Permute effective weights / biases given in the reference's order into packed order."""
        device = weights[0].device
        perms = self.tables(device)[2]
        pw, pb = [], []
        for l, (w, b) in enumerate(zip(weights, biases)):
            if l == self.n_layers - 1 and self.has_padding:
                # -1 entries select an appended zero row
                w = torch.cat([w, w.new_zeros(1, w.shape[1])])
                b = torch.cat([b, b.new_zeros(1)])
            w = w.index_select(0, perms[l + 1])
            if l > 0:
                w = w.index_select(1, perms[l])
            # rows padded to 16 bytes (zero columns): 16-byte aligned operands for the 128 x 128 GEMM kernel
            pad = (-w.shape[1]) % (16 // w.element_size())
            if pad:
                w = torch.nn.functional.pad(w, (0, pad))[:, :w.shape[1]]
            else:
                w = w.contiguous()
            pw.append(w)
            pb.append(b.index_select(0, perms[l + 1]).contiguous())
        return pw, pb

    def degree_rows(self, layer, degree):
        """[begin, end) of the packed units of ``layer`` (1-based: 1..n_layers) that have ``degree``.
        Only meaningful for degree-sorted layers."""
        d = self.packed_degrees[layer]
        idx = (d == degree).nonzero().flatten()
        if len(idx) == 0:
            return 0, 0
        return int(idx[0]), int(idx[-1]) + 1

    def degree_prefix(self, layer, degree, strict):
        """Number of packed units of ``layer`` (0 = input, must be sorted) with degree < (or <=) ``degree``."""
        d = self.packed_degrees[layer]
        return int(((d < degree) if strict else (d <= degree)).sum())
