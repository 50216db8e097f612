"""Flow wrapper constraining the (weighted) centroid (reference tfep/nn/flows/centroid.py:30-263)."""

import torch

from ... import _frames
from ...utils.geometry import atom_to_flattened, atom_to_flattened_indices, flattened_to_atom
from ...utils.misc import ensure_tensor_sequence
from .partial import PartialFlow


class CenteredCentroidFlow(PartialFlow):
    """Translate the centroid of the points to ``origin``, run the wrapped flow on all coordinates except those of one
    fixed point, then place the fixed point so that the centroid is preserved (and optionally translate back).

    The features are ``n_points`` points of ``space_dimension`` coordinates each, point after point.  Constructor
    arguments, defaults and error messages as in the reference (centroid.py:107-160): ``subset_point_indices`` /
    ``weights`` select and weight the points that define the centroid, ``fixed_point_idx`` is relative to the subset.
    """

    def __init__(self, flow, space_dimension, subset_point_indices=None, weights=None, fixed_point_idx=0, origin=None,
                 translate_back=True, return_partial=False):
        if return_partial and translate_back:
            raise ValueError("'return_partial=True' is supported only if 'translate_back=False'")
        if origin is None:
            origin = torch.zeros(space_dimension)
        else:
            if len(origin) != space_dimension:
                raise ValueError("'origin' must have length equal to 'space_dimension'.")
            origin = ensure_tensor_sequence(origin)
        if subset_point_indices is not None:
            subset_point_indices = ensure_tensor_sequence(subset_point_indices)
        if weights is not None:
            weights = ensure_tensor_sequence(weights)
        if subset_point_indices is None:
            fixed_point = fixed_point_idx
        else:
            fixed_point = subset_point_indices[fixed_point_idx]
            if weights is not None and len(weights) != len(subset_point_indices):
                raise ValueError("'weights' must have the same length as 'subset_point_indices'.")
        super().__init__(flow, fixed_indices=atom_to_flattened_indices(torch.tensor([int(fixed_point)]), space_dimension),
                         return_partial=return_partial)
        if weights is not None:
            weights = (weights / torch.sum(weights)).unsqueeze(dim=1)
        self._space_dimension = space_dimension
        self.register_buffer('_fixed_point_idx', torch.as_tensor(fixed_point_idx))
        self.register_buffer('_subset_point_indices', subset_point_indices)
        self.register_buffer('_weights', weights)
        self.register_buffer('origin', origin)
        self.translate_back = translate_back
        self._frame = None

    @property
    def space_dimension(self):
        return self._space_dimension

    def forward(self, x):
        return self._transform(x)

    def inverse(self, y):
        if not self.translate_back:
            raise ValueError("The inverse of CenteredCentroidFlow can be computed"
                             " only if 'translate_back' is set to True during both"
                             " the forward and inverse transformations.")
        return self._transform(y, inverse=True)

    def _centroid(self, points, exclude_fixed_point=False):
        """(batch, dim) centroid of the defining points; with ``exclude_fixed_point`` the fixed point's own term is
        left out and its weight is returned as well."""
        if self._subset_point_indices is not None:
            points = points[:, self._subset_point_indices]
        if self._weights is None:
            centroid = points.mean(dim=1)
            fixed_weight = 1 / points.shape[1]
        else:
            centroid = (points * self._weights).sum(dim=1)
            fixed_weight = self._weights[self._fixed_point_idx]
        if exclude_fixed_point:
            return centroid - points[:, self._fixed_point_idx] * fixed_weight, fixed_weight
        return centroid

    def _fused(self, x, inverse):
        """Inference on the GPU: one fused kernel each side of the wrapped flow (tfep_b200/csrc/frames.cu)."""
        if self._frame is None or self._frame.n_features != x.shape[1]:
            subset = None if self._subset_point_indices is None else self._subset_point_indices.tolist()
            slot = int(self._fixed_point_idx)
            self._frame = _frames.CentroidFrame(
                x.shape[1], self._space_dimension, self._fixed_indices.tolist(), subset,
                None if self._weights is None else self._weights.detach().cpu(), slot,
                slot if subset is None else subset[slot], self.origin.detach().cpu().tolist(),
                restore=subset is None or len(subset) > 1, translate_back=self.translate_back)
        x, x_prop, shift = self._frame.pre(x)
        out = self.flow.inverse(x_prop) if inverse else self.flow(x_prop)
        if self.return_partial:
            return out
        return (self._frame.post(x, out[0], shift), *out[1:])

    def _transform(self, x, inverse=False):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_grad and _frames.usable(x) and self._space_dimension <= 4:
            return self._fused(x, inverse)
        dim = self._space_dimension
        points = flattened_to_atom(x, dim)
        shift = (self.origin - self._centroid(points)).unsqueeze(dim=1)
        x_centered = atom_to_flattened(points + shift)
        y, log_det_J = super().inverse(x_centered) if inverse else super().forward(x_centered)
        if self.return_partial:
            return y, log_det_J
        if self._subset_point_indices is None or len(self._subset_point_indices) > 1:
            rest, fixed_weight = self._centroid(flattened_to_atom(y, dim), exclude_fixed_point=True)
            y = y.index_copy(1, self._fixed_indices, (self.origin - rest) / fixed_weight)
        if self.translate_back:
            y = atom_to_flattened(flattened_to_atom(y, dim) - shift)
        return y, log_det_J
