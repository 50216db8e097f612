"""Flow wrapper that keeps some degrees of freedom fixed (reference tfep/nn/flows/partial.py:25-121)."""

import torch

from ...utils.misc import ensure_tensor_sequence


class PartialFlow(torch.nn.Module):
    """Run ``flow`` on all features except ``fixed_indices``, which are copied through unchanged.

    The wrapped flow sees a contiguous ``(batch, n_features - n_fixed)`` tensor; unlike conditioning features of a
    ``MAF`` (degree -1) the fixed features are invisible to it.

    Parameters
    ----------
    flow : torch.nn.Module
    fixed_indices : sequence of int
    return_partial : bool
        If True only the propagated features are returned.
    """

    def __init__(self, flow, fixed_indices, return_partial=False):
        super().__init__()
        self.flow = flow
        self.return_partial = return_partial
        self.register_buffer('_fixed_indices', ensure_tensor_sequence(fixed_indices))
        self.register_buffer('_propagated_indices', None)        # built lazily: needs the input width

    def n_parameters(self):
        return self.flow.n_parameters()

    def forward(self, x):
        return self._pass(x, inverse=False)

    def inverse(self, y):
        return self._pass(y, inverse=True)

    def _propagated(self, n_features, device):
        if self._propagated_indices is None or self._propagated_indices.device != device:
            keep = torch.ones(n_features, dtype=torch.bool)
            keep[self._fixed_indices.cpu()] = False
            self._propagated_indices = keep.nonzero().flatten().to(device)
        return self._propagated_indices

    def _pass(self, x, inverse):
        n_fixed = len(self._fixed_indices)
        full = x
        if n_fixed > 0:
            prop = self._propagated(x.shape[1], x.device)
            x = x.index_select(1, prop)
        out = self.flow.inverse(x) if inverse else self.flow(x)
        if self.return_partial:
            return out
        if n_fixed == 0:
            return out
        # out-of-place scatter: fixed columns keep their input values, the others take the flow's output
        y = full.index_copy(1, prop, out[0])
        return (y, *out[1:])
