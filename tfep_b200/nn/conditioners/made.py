"""MADE conditioner (Masked Autoencoder for Distribution Estimation) on the sm_100a kernels.

Public surface of the reference's ``tfep/nn/conditioners/made.py``: ``generate_degrees`` (:32-145) and
``MADE`` (:152-434) with the same constructor arguments, ``layers`` container (even entries are the
masked linear layers holding ``weight_g / weight_v / bias / mask``, odd entries ``ELU``), properties and
``set_output``.  ``forward`` does not walk the container: it packs the effective weights in degree-sorted
order (tfep_b200/_pack.py) and runs all layers as one autograd node with ELU fused into the GEMM epilogue.
"""

from collections.abc import Sequence
from typing import Literal, Optional, Union

import numpy as np
import torch

from ... import _ops
from ..._pack import MadePlan
from ...utils.misc import ensure_tensor_sequence
from . import conditioner as _conditioner
from .. import masked


def generate_degrees(
        n_features: int,
        order: Literal['ascending', 'descending', 'random'] = 'ascending',
        max_value: Optional[int] = None,
        conditioning_indices: Optional[Sequence[int]] = None,
        repeats: Union[int, Sequence[int]] = 1,
) -> torch.Tensor:
    """Generate node degrees for MADE layers (reference nn/conditioners/made.py:32-145).

    Degrees run from 0 to ``max_value`` (tiled if there are fewer values than features), optionally
    repeated ``repeats`` times each; conditioning features get degree -1.

    >>> generate_degrees(7, order='descending', max_value=2).tolist()
    [2, 1, 0, 2, 1, 0, 2]
    >>> generate_degrees(7, repeats=[1, 3, 2], conditioning_indices=[2]).tolist()
    [0, 1, -1, 1, 1, 2, 2]
    """
    n_free = n_features if conditioning_indices is None else n_features - len(conditioning_indices)
    if max_value is None:
        try:
            max_value = len(repeats) - 1
        except TypeError:
            max_value = int(np.ceil(n_free / repeats)) - 1
    if order == 'ascending':
        degrees = torch.arange(max_value + 1)
    elif order == 'descending':
        degrees = torch.arange(max_value, -1, -1)
    elif order == 'random':
        degrees = torch.randperm(max_value + 1)
    else:
        raise ValueError("Accepted string values for 'order' are 'ascending', 'descending', and 'random'.")
    repeats = ensure_tensor_sequence(repeats, dtype=int)
    degrees = _round_robin(torch.repeat_interleave(degrees, repeats)[:n_free], length=n_free)
    if conditioning_indices is None:
        return degrees
    try:
        conditioning_indices = conditioning_indices.tolist()
    except AttributeError:
        pass
    taken = set(conditioning_indices)
    free = [i for i in range(n_features) if i not in taken]
    out = torch.empty(n_features, dtype=degrees.dtype)
    out[list(conditioning_indices)] = -1
    out[free] = degrees
    return out


class MADE(_conditioner.Conditioner):
    """Autoregressive conditioner made of masked linear layers with ELU in between.

    An output node of degree ``i`` depends only on inputs of degree strictly less than ``i``; inputs of
    degree -1 ("conditioning") feed every output.  See the reference docstring (made.py:152-244).

    Parameters
    ----------
    degrees_in, degrees_out : Sequence[int]
        Degrees of the input / output nodes.
    hidden_layers : int, Sequence[int] or Sequence[Sequence[int]]
        Number of hidden layers (default width ``max(ceil(sqrt(n_in_relevant * n_out)), n_in_relevant)``),
        or their widths (degrees assigned round-robin), or their degrees.
    weight_norm : bool
        Apply (masked) weight normalisation to every linear layer.
    """

    def __init__(self, degrees_in, degrees_out, hidden_layers=2, weight_norm=True):
        super().__init__()
        degrees_in = ensure_tensor_sequence(degrees_in, dtype=int)
        degrees_out = ensure_tensor_sequence(degrees_out, dtype=int)
        degrees_hidden = self._get_degrees_hidden(degrees_in, degrees_out, hidden_layers)
        chain = [degrees_in] + list(degrees_hidden) + [degrees_out]

        layers = []
        for l in range(len(chain) - 1):
            is_output = l == len(chain) - 2
            mask = masked.create_autoregressive_mask(chain[l], chain[l + 1], strictly_less=is_output, transpose=True)
            lin = masked.MaskedLinear(len(chain[l]), len(chain[l + 1]), bias=True, mask=mask)
            if weight_norm:
                lin = masked.masked_weight_norm(lin, name='weight')
            layers.extend([lin, torch.nn.ELU()])
        layers.pop()
        self.layers = torch.nn.Sequential(*layers)

        self._degree_chain = [d.clone() for d in chain]
        self._plan = MadePlan(self._degree_chain)       # hidden units degree-sorted, output in reference order
        self._packed_cache = {}
        self._packed_epoch = 0
        # loading a checkpoint writes the parameters in place: drop every packed copy derived from the old values
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate_packed())

    # -- reference properties -------------------------------------------------------------------
    @property
    def dimension_in(self) -> int:
        return self.layers[0].in_features

    @property
    def dimension_out(self) -> int:
        return self.layers[-1].out_features

    @property
    def dimensions_hidden(self) -> torch.Tensor:
        return torch.tensor([l.out_features for l in self.layers[:-1:2]])

    @property
    def weight_norm(self):
        return hasattr(self.layers[-1], 'weight_g')

    def n_parameters(self) -> int:
        """The total number of (unmasked) parameters."""
        return sum(l.n_parameters() for l in self.layers[::2])

    def set_output(self, output: torch.Tensor):
        """Make the conditioner return ``output`` for any input (reference made.py:358-364)."""
        last = self.layers[-1]
        with torch.no_grad():
            (last.weight_g if self.weight_norm else last.weight).zero_()
        last.bias.data = output.to(last.bias.data)
        self.invalidate_packed()

    # -- evaluation -----------------------------------------------------------------------------
    def _linear_layers(self):
        return list(self.layers[::2])

    def effective_weights(self):
        """[(W_eff, bias)] in the reference's unit order, differentiable w.r.t. the parameters."""
        out = []
        for lin in self._linear_layers():
            if hasattr(lin, 'weight_g'):
                w = masked.effective_weight(lin.weight_v, lin.weight_g, lin.mask)
            else:
                w = lin.weight * lin.mask
            out.append((w, lin.bias))
        return out

    def invalidate_packed(self):
        """Forget every packed / padded / bf16 copy of the weights (this module's, the fused kernels' and the inverse
        sweep's: they all key on :meth:`_param_versions`).  The caches notice ordinary in-place updates (optimizer steps,
        ``copy_`` on the parameter, ``load_state_dict``) through the tensors' version counters; writes through ``.data``
        (``p.data.copy_(ema)``, ``p.data.clamp_()``) bypass those counters and MUST be followed by this call."""
        self._packed_epoch += 1
        self._packed_cache.clear()
        self.__dict__.pop('_key_tensors', None)

    def _param_versions(self):
        """Cache key of everything derived from the weights: parameters AND mask buffers (pointer, version) plus the
        explicit invalidation epoch."""
        tensors = self.__dict__.get('_key_tensors')
        if tensors is None:         # walked once: the module tree does not change after construction
            tensors = list(self.parameters()) + [lin.mask for lin in self._linear_layers()]
            self.__dict__['_key_tensors'] = tensors
        return (self._packed_epoch, *[(t.data_ptr(), t._version) for t in tensors])

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .double() may REPLACE parameter and buffer tensors: forget the cached list of key tensors
        self.__dict__.pop('_key_tensors', None)
        out = super()._apply(fn, *args, **kwargs)
        self.__dict__.pop('_key_tensors', None)
        return out

    def packed_weights(self, plan=None):
        """Packed (degree-sorted) effective weights for ``plan``; cached while gradients are off."""
        plan = self._plan if plan is None else plan
        track = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if track:
            lins = self._linear_layers()
            if all(hasattr(l, 'weight_g') and l.weight_v.is_cuda and l.weight_v.dtype == torch.float32 and l.bias is not None
                   for l in lins):
                # training step: normalisation, mask and packing of a layer in one launch, their VJP in another
                # (tfepb_wn_pack) instead of ~35 tensor-algebra launches per layer
                perms = plan.perm_tables(lins[0].weight_v.device)
                pw, pb = [], []
                for l, lin in enumerate(lins):
                    w, b = _ops.wn_pack(lin.weight_v, lin.weight_g, lin.bias, lin.mask, perms[l + 1], perms[l])
                    pw.append(w)
                    pb.append(b)
                return pw, pb
            ws, bs = zip(*self.effective_weights())
            return plan.pack(list(ws), list(bs))
        key = (id(plan), self._param_versions())
        hit = self._packed_cache.get(id(plan))
        if hit is None or hit[0] != key:
            with torch.no_grad():
                ws, bs = zip(*self.effective_weights())
                hit = (key, plan.pack(list(ws), list(bs)))
            self._packed_cache[id(plan)] = hit
        return hit[1]

    def run_plan(self, x, plan, precision='fp32'):
        """Conditioner output in the packed output order of ``plan``.  ``precision='bf16'``: the three products of
        every layer (forward, and both backward products under autograd) run on the tensor cores."""
        pw, pb = self.packed_weights(plan)
        if precision in ('bf16', 'bf16x3', 'bf16x6'):
            if x.dtype != torch.float32:
                raise _ops._lib.TfepB200Error(f"precision={precision!r} takes float32 inputs")
            kb_fwd, kb_bwd, rr_w = plan.tc_ranges(x.device)
            n_split = {'bf16': 1, 'bf16x3': 2, 'bf16x6': 3}[precision]
            images = None
            if n_split > 1 and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
                # split images of the packed weights, cached like the packed weights themselves
                key = ('tc_images', id(plan), n_split, self._param_versions())
                hit = self._packed_cache.get(key[:3])
                if hit is None or hit[0] != key:
                    hit = (key, [_ops.tc_pack(w, 256, n_split=n_split) for w in pw])
                    self._packed_cache[key[:3]] = hit
                images = hit[1]
            return _ops.made_forward_tc(x, pw, pb, kb_fwd, kb_bwd, rr_w, n_split=n_split, weight_images=images)
        k_ranges, n_ranges, _ = plan.tables(x.device)
        return _ops.made_forward(x, pw, pb, k_ranges, n_ranges)

    def forward(self, x):
        return self.run_plan(x, self._plan)

    # -- degrees --------------------------------------------------------------------------------
    @classmethod
    def _get_degrees_hidden(cls, degrees_in, degrees_out, hidden_layers):
        """Degrees of the hidden layers (reference made.py:366-434)."""
        try:
            hidden_layers = hidden_layers.tolist()
        except AttributeError:
            pass
        max_degree_out = degrees_out.max()
        relevant = degrees_in < max_degree_out
        if isinstance(hidden_layers, int):
            n_relevant = relevant.sum().tolist()
            width = max(int(np.ceil((n_relevant * len(degrees_out)) ** 0.5)), n_relevant)
            hidden_layers = [width for _ in range(hidden_layers)]
        if isinstance(hidden_layers[0], int):
            motif = degrees_in[relevant]
            return [_round_robin(motif, length=width, err_msg=(
                f'Hidden layer {idx} is too small for the number of input features. Increase the size of the '
                'layer or explicitly pass the degrees for the hidden layers.'))
                for idx, width in enumerate(hidden_layers)]
        degrees_hidden = [ensure_tensor_sequence(h) for h in hidden_layers]
        for idx, degrees in enumerate(degrees_hidden):
            if torch.any(degrees >= max_degree_out):
                raise ValueError(f'The {idx}-th hidden layer contain nodes with degrees that will be ignored '
                                 'by the output layer.')
        return degrees_hidden


def _round_robin(x: torch.Tensor, length: int, err_msg: Optional[str] = None) -> torch.Tensor:
    """Tile ``x`` until ``length`` elements are produced (reference made.py:441-461)."""
    n_tiles, n_rest = divmod(length, len(x))
    if n_tiles == 0:
        raise ValueError(err_msg if err_msg is not None else f'Length {length} is smaller than the array (len={len(x)}).')
    out = torch.tile(x, (n_tiles,))
    if n_rest != 0:
        out = torch.cat([out, x[:n_rest]])
    return out
