"""Flow wrapper constraining the orientation of the frame of reference (reference tfep/nn/flows/oriented.py:38-225)."""

import torch

from ... import _frames
from ...utils.geometry import (atom_to_flattened, atom_to_flattened_indices, batchwise_rotate, flattened_to_atom,
                               get_axis_from_name, reference_frame_rotation_matrix)
from .partial import PartialFlow


class OrientedFlow(PartialFlow):
    """Rotate every sample so that one point lies on ``axis`` and a second one on ``plane``, run the wrapped flow on
    the remaining degrees of freedom (the three constrained coordinates are removed), and optionally rotate back.
    3D only.  Constructor arguments, defaults and error messages as in the reference (oriented.py:66-160).
    """

    def __init__(self, flow, axis_point_idx=None, plane_point_idx=None, axis='x', plane='xy',
                 round_off_imprecisions=True, rotate_back=True, return_partial=False):
        if return_partial and rotate_back:
            raise ValueError("'return_partial=True' is supported only if 'rotate_back=False'")
        if axis_point_idx is None:
            axis_point_idx = 0 if plane_point_idx != 0 else 1
        if plane_point_idx is None:
            plane_point_idx = 0 if axis_point_idx != 0 else 1
        if axis_point_idx == plane_point_idx:
            raise ValueError("'axis_point_idx' and 'plane_point_idx' must be different.")
        if axis not in plane:
            raise ValueError("To constrain 'plane_atom_idx' to stay on plane {plane} "
                             "'axis_atom_idx' must be constrained on an axis on the same plane.")
        axis_vector = get_axis_from_name(axis)
        plane_axis_vector = get_axis_from_name([n for n in 'xyz' if n != axis and n in plane][0])
        plane_normal_vector = torch.linalg.cross(axis_vector, plane_axis_vector)
        # two coordinates of the point on the axis and the out-of-plane coordinate of the second point are constrained
        fixed_indices = torch.cat([atom_to_flattened_indices(torch.tensor([axis_point_idx]))[axis_vector == 0.0],
                                   atom_to_flattened_indices(torch.tensor([plane_point_idx]))[plane_normal_vector != 0.0]])
        super().__init__(flow, fixed_indices=fixed_indices, return_partial=return_partial)
        self.register_buffer('_axis', axis_vector)
        self.register_buffer('_plane_axis', plane_axis_vector)
        self.register_buffer('_plane_normal', plane_normal_vector)
        self.register_buffer('_axis_point_idx', torch.as_tensor(axis_point_idx))
        self.register_buffer('_plane_point_idx', torch.as_tensor(plane_point_idx))
        self.round_off_imprecisions = round_off_imprecisions
        self.rotate_back = rotate_back
        self._frame = None

    def forward(self, x):
        return self._transform(x)

    def inverse(self, y):
        if not self.rotate_back:
            raise ValueError("The inverse of OrientedFlow can be computed only"
                             " if 'rotate_back' is set to True during both the"
                             " forward and inverse transformations.")
        return self._transform(y, inverse=True)

    def _fused(self, x, inverse):
        """Inference on the GPU: one fused kernel each side of the wrapped flow (tfep_b200/csrc/frames.cu)."""
        if self._frame is None or self._frame.n_features != x.shape[1]:
            self._frame = _frames.OrientedFrame(
                x.shape[1], self._fixed_indices.tolist(), int(self._axis_point_idx), int(self._plane_point_idx),
                int(self._axis.argmax()), int(self._plane_axis.argmax()), self.round_off_imprecisions, self.rotate_back)
        x, x_prop, rot = self._frame.pre(x)
        out = self.flow.inverse(x_prop) if inverse else self.flow(x_prop)
        if self.return_partial:
            return out
        return (self._frame.post(x, out[0], rot), *out[1:])

    def _transform(self, x, inverse=False):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_grad and _frames.usable(x) and x.shape[1] % 3 == 0:
            return self._fused(x, inverse)
        points = flattened_to_atom(x)
        rotations = reference_frame_rotation_matrix(
            axis_atom_positions=points[:, self._axis_point_idx], plane_atom_positions=points[:, self._plane_point_idx],
            axis=self._axis.to(x), plane_axis=self._plane_axis.to(x), plane_normal=self._plane_normal.to(x),
            project_on_positive_axis=False)
        x = atom_to_flattened(batchwise_rotate(points, rotations))
        if self.round_off_imprecisions:
            x = x.index_fill(1, self._fixed_indices, 0.0)          # exactly zero by construction
        y, log_det_J = super().inverse(x) if inverse else super().forward(x)
        if self.return_partial:
            return y, log_det_J
        if self.rotate_back:
            y = atom_to_flattened(batchwise_rotate(flattened_to_atom(y), rotations, inverse=True))
        return y, log_det_J
