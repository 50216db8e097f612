"""Transformer applying different transformers to disjoint groups of features.

Reference: tfep/nn/transformers/mixed.py:29-186.  Here the children are lowered into ONE kernel program
(a list of launches sharing x, y, the parameter matrix and the log-det accumulator), so the mixture costs
no gather / scatter of feature columns.
"""

from collections.abc import Sequence

import torch

from ...utils.misc import ensure_tensor_sequence
from .transformer import MAFTransformer


class MixedTransformer(MAFTransformer):
    """Mix transformers over feature groups: ``indices[i]`` are the input features of ``transformers[i]``.

    Parameters are expected grouped by transformer (all parameters of the first transformer, then those of
    the second, ...), each group in that transformer's own layout.
    """

    def __init__(self, transformers: Sequence[MAFTransformer], indices: Sequence[Sequence[int]]):
        super().__init__()
        if len(transformers) < 2:
            raise ValueError('The number of transformers must be greater than 1.')
        if len(transformers) != len(indices):
            raise ValueError('The number of elements in indices must equal that in transformers.')
        self._transformers = transformers      # plain list, like the reference (mixed.py:58)
        for idx, ind in enumerate(indices):
            self.register_buffer(f'_indices{idx}', ensure_tensor_sequence(ind))
        par_lengths = [len(t.get_identity_parameters(len(ind))) for t, ind in zip(transformers, indices)]
        self.register_buffer('_parameters_split_indices', torch.cumsum(torch.tensor(par_lengths[:-1]), dim=0))
        self._par_offsets = [0] + torch.cumsum(torch.tensor(par_lengths), dim=0).tolist()[:-1]
        self._host_indices = [ensure_tensor_sequence(ind).long().cpu().clone() for ind in indices]

    @property
    def _indices(self):
        return [getattr(self, f'_indices{idx}') for idx in range(len(self._transformers))]

    def _apply(self, fn, *args, **kwargs):
        # children live in a plain list (reference behaviour): move / cast them along explicitly
        for t in self._transformers:
            t._apply(fn, *args, **kwargs)
        return super()._apply(fn, *args, **kwargs)

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        return torch.cat([t.get_identity_parameters(len(ind)) for t, ind in zip(self._transformers, self._host_indices)],
                         dim=-1)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return torch.cat([t.get_degrees_out(degrees_in[ind]) for t, ind in zip(self._transformers, self._host_indices)],
                         dim=-1)

    def _parts(self, n_features):
        parts = []
        for t, ind, off in zip(self._transformers, self._host_indices, self._par_offsets):
            parts.extend(p.moved(off, ind) for p in t._parts(len(ind)))
        return parts
