"""Embedding layers for MAF conditioners (reference tfep/nn/embeddings/mafembed.py:30-172).

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
        lim = self.limits.detach().double().cpu()
        self._lower = float(lim[0])
        self._scale = float(2 * math.pi / (lim[1] - lim[0]))
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

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = tuple(t.to(device) for t in self._host_tables)
        return self._dev[key]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _PeriodicEmbeddingFunction.apply(x, self)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return torch.cat([degrees_in[self._nonperiodic_indices],
                          degrees_in[self._periodic_indices].repeat_interleave(2)])
