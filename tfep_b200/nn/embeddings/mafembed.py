"""Embedding layers for MAF conditioners (reference tfep/nn/embeddings/mafembed.py:30-172).

``FlipInvariantEmbedding`` and ``MixedEmbedding`` (reference mafembed.py:174-446) are small learnable / composite
modules evaluated with tensor algebra in front of the conditioner kernels.

``PeriodicEmbedding`` lifts periodic features to ``(cos, sin)`` before they enter the MADE conditioner -- what the
reference's ``MixedMAFMap`` runs (app/mixedmaf.py:341-353).  Kernels: tfepb_periodic_embedding /
tfepb_periodic_embedding_backward (one thread per input element, forward plus hand-written backward); inside
``MAF.inverse`` the lift is applied by the persistent sweep kernel itself as each feature is inverted.
"""

import abc
import ctypes
import math

import torch

from ... import _lib
from ..._lib import check, dtype_code, ptr, stream_ptr
from ...utils.misc import ensure_tensor_sequence


class MAFEmbedding(abc.ABC, torch.nn.Module):
    """An embedding layer compatible with :class:`tfep_b200.nn.flows.MAF`."""

    @abc.abstractmethod
    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        """Degrees of the features after the forward pass (the conditioner's input degrees)."""


class _PeriodicEmbeddingFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, emb):
        _lib.require_cuda(x)
        x = x.contiguous()
        B, n_in = x.shape
        out_col, periodic = emb._tables(x.device)
        out = torch.empty(B, emb.n_features_out, dtype=x.dtype, device=x.device)
        if B > 0:
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_periodic_embedding(dtype_code(x), ptr(x), n_in, B, n_in, ptr(out_col), ptr(periodic),
                                                           emb._lower, emb._scale, ptr(out), emb.n_features_out,
                                                           stream_ptr(x)))
        ctx.save_for_backward(x)
        ctx.emb = emb
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        emb = ctx.emb
        grad_out = grad_out.contiguous()
        B, n_in = x.shape
        out_col, periodic = emb._tables(x.device)
        gx = torch.empty_like(x)
        if B > 0:
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_periodic_embedding_backward(
                    dtype_code(x), ptr(x), n_in, B, n_in, ptr(out_col), ptr(periodic), emb._lower, emb._scale,
                    ptr(grad_out), emb.n_features_out, ptr(gx), n_in, stream_ptr(x)))
        return gx, None


class PeriodicEmbedding(MAFEmbedding):
    """Lift periodic degrees of freedom into a periodic representation ``(cos, sin)``.

    The output holds the non-periodic features first (in order) and then one ``(cos, sin)`` pair per periodic
    feature (reference mafembed.py:128-142); both members of a pair get the degree of their feature.

    Parameters
    ----------
    n_features_in : int
    limits : Sequence[float]
        ``(lower, upper)`` of the periodic variables; the period is ``upper - lower``.
    periodic_indices : Sequence[int], optional
        Ordered indices of the periodic input features.  Default: all features.
    """

    def __init__(self, n_features_in, limits, periodic_indices=None):
        super().__init__()
        self.register_buffer('limits', ensure_tensor_sequence(limits))
        if periodic_indices is None:
            periodic_indices = torch.arange(n_features_in)
        else:
            periodic_indices = ensure_tensor_sequence(periodic_indices)
            if len(periodic_indices.unique()) < len(periodic_indices):
                raise ValueError('Found duplicated indices in periodic_indices.')
        self.register_buffer('_periodic_indices', periodic_indices)
        all_idx = torch.arange(n_features_in)
        self.register_buffer('_nonperiodic_indices', all_idx[~torch.isin(all_idx, periodic_indices)])
        self.n_features_in = int(n_features_in)
        self.n_features_out = int(n_features_in + len(periodic_indices))
        self._refresh_limits()
        # `limits` is a buffer: a checkpoint built with other limits must change the constants the kernels receive
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._refresh_limits())
        # per input column: first output column and whether it is lifted
        out_col = torch.empty(n_features_in, dtype=torch.int32)
        periodic = torch.zeros(n_features_in, dtype=torch.int32)
        nonper = self._nonperiodic_indices.tolist()
        for j, c in enumerate(nonper):
            out_col[c] = j
        for j, c in enumerate(self._periodic_indices.tolist()):
            out_col[c] = len(nonper) + 2 * j
            periodic[c] = 1
        self._host_tables = (out_col, periodic)
        self._dev = {}

    def _refresh_limits(self):
        """Kernel constants of the lift, from the `limits` buffer (host floats: read once per change, not per call)."""
        lim = self.limits.detach().double().cpu()
        self._lower = float(lim[0])
        self._scale = float(2 * math.pi / (lim[1] - lim[0]))

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = tuple(t.to(device) for t in self._host_tables)
        return self._dev[key]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _PeriodicEmbeddingFunction.apply(x, self)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        dev = degrees_in.device                  # (the degrees usually live on the host, the module may not)
        return torch.cat([degrees_in[self._nonperiodic_indices.to(dev)],
                          degrees_in[self._periodic_indices.to(dev)].repeat_interleave(2)])


def _complement(n_features_in, indices):
    """Sorted indices of range(n_features_in) that are not in ``indices``."""
    keep = torch.ones(n_features_in, dtype=torch.bool)
    keep[indices] = False
    return keep.nonzero().flatten()


class FlipInvariantEmbedding(MAFEmbedding):
    """Embed vector features (e.g. quaternions) into a representation invariant to a sign flip of the vector
    (reference mafembed.py:174-351, Koehler et al. 2023, SI eq. 46): a softmax-weighted average of a small
    network evaluated at ``v`` and ``-v``, ``sum_s softmax_s(w(s v)) e(s v)``.

    Output: the non-embedded features first, then ``embedding_dimension`` values per vector; all components of a
    vector must share one degree.  Parameter names (``embedding_layer.{0,2}``, ``weight_layer.{0,2}``) as in the reference.
    """

    def __init__(self, n_features_in, embedding_dimension, embedded_indices=None, vector_dimension=4, hidden_layer_width=32):
        super().__init__()
        self.embedding_layer = torch.nn.Sequential(
            torch.nn.Linear(vector_dimension, hidden_layer_width), torch.nn.ELU(),
            torch.nn.Linear(hidden_layer_width, embedding_dimension))
        self.weight_layer = torch.nn.Sequential(
            torch.nn.Linear(vector_dimension, hidden_layer_width), torch.nn.ELU(),
            torch.nn.Linear(hidden_layer_width, 1))
        if embedded_indices is None:
            embedded_indices = torch.arange(n_features_in)
        else:
            embedded_indices = ensure_tensor_sequence(embedded_indices)
            if len(embedded_indices.unique()) < len(embedded_indices):
                raise ValueError('Found duplicated indices in embedded_indices.')
        self.register_buffer('_embedded_indices', embedded_indices)
        self.register_buffer('_nonembedded_indices', _complement(n_features_in, embedded_indices))

    @property
    def vector_dimension(self) -> int:
        return self.embedding_layer[0].in_features

    @property
    def embedding_dimension(self) -> int:
        return self.embedding_layer[-1].out_features

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        batch = x.shape[0]
        v = x[:, self._embedded_indices].reshape(-1, self.vector_dimension)
        both = torch.stack([v, -v], dim=1)                                  # (n_vectors_total, 2, d)
        weights = torch.softmax(self.weight_layer(both), dim=1)             # (n_vectors_total, 2, 1)
        embedded = (weights * self.embedding_layer(both)).sum(dim=1)
        return torch.cat([x[:, self._nonembedded_indices], embedded.reshape(batch, -1)], dim=1)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        dev = degrees_in.device
        vec = degrees_in[self._embedded_indices.to(dev)].reshape(-1, self.vector_dimension)
        if not torch.all(vec == vec[:, [0]]):
            raise ValueError('The same degree must be assigned to all '
                             'components of each embedded vectors.')
        return torch.cat([degrees_in[self._nonembedded_indices.to(dev)],
                          vec[:, [0]].expand(-1, self.embedding_dimension).flatten()])


class MixedEmbedding(MAFEmbedding):
    """Apply several embedding layers, each to its own subset of the features (reference mafembed.py:354-446).
    Output: the features no layer takes first, then the outputs of the layers in order."""

    def __init__(self, n_features_in, embedding_layers, embedded_indices):
        super().__init__()
        if len(embedding_layers) != len(embedded_indices):
            raise ValueError('Different number of layers and indices.')
        embedded_indices = [ensure_tensor_sequence(i) for i in embedded_indices]
        first = set(embedded_indices[0].tolist())
        for indices in embedded_indices[1:]:
            if len(first & set(indices.tolist())) > 0:
                raise ValueError('Different embedding layers must be assigned '
                                 'to different feature indices.')
        self.embedding_layers = torch.nn.ModuleList(embedding_layers)
        for i, indices in enumerate(embedded_indices):
            self.register_buffer(f'_embedded_indices{i}', indices)
        self.register_buffer('_nonembedded_indices', _complement(n_features_in, torch.cat(embedded_indices)))

    def _indices(self, i):
        return getattr(self, f'_embedded_indices{i}')

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        parts = [layer(x[:, self._indices(i)].contiguous()) for i, layer in enumerate(self.embedding_layers)]
        return torch.cat([x[:, self._nonembedded_indices], *parts], dim=1)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        dev = degrees_in.device
        parts = [layer.get_degrees_out(degrees_in[self._indices(i).to(dev)]) for i, layer in enumerate(self.embedding_layers)]
        return torch.cat([degrees_in[self._nonembedded_indices.to(dev)], *parts])
