"""Affine transformer ``y = x exp(a) + b`` and volume-preserving shift ``y = x + b``
(reference tfep/nn/transformers/affine.py:28-141, 148-275, 281-456).

Kernels: tfepb_affine / tfepb_affine_backward, tfepb_shift / tfepb_shift_backward.
"""

import torch

from ... import _program
from .transformer import MAFTransformer


class AffineTransformer(MAFTransformer):
    r"""Affine transformer :math:`y_i = \exp(a_i) x_i + b_i`.

    ``parameters[:, i]`` is the shift :math:`b_i` and ``parameters[:, n_features + i]`` the log-scale
    :math:`a_i` of feature ``i``; ``log_det_J = sum_i a_i``.
    """

    n_parameters_per_feature = 2

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        return torch.zeros(size=(self.n_parameters_per_feature * n_features,))

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.tile((self.n_parameters_per_feature,))

    def _parts(self, n_features):
        return [_program.Part('affine', None, n_features, 2)]


class _ShiftSpec:
    """Per-feature period / lower-limit tables of a volume-preserving shift (0 = not periodic)."""

    def __init__(self, n_features, periodic_indices, periodic_limits):
        self.period = torch.zeros(n_features, dtype=torch.float64)
        self.lower = torch.zeros(n_features, dtype=torch.float64)
        if periodic_indices is not None:
            idx = torch.as_tensor(periodic_indices).long().cpu()
            lim = torch.as_tensor(periodic_limits).double().cpu()
            self.period[idx] = lim[1] - lim[0]
            self.lower[idx] = lim[0]
        self._dev = {}

    def tables(self, dtype, device):
        key = (dtype, str(device))
        if key not in self._dev:
            self._dev[key] = (self.period.to(device=device, dtype=dtype), self.lower.to(device=device, dtype=dtype))
        return self._dev[key]


class VolumePreservingShiftTransformer(MAFTransformer):
    r"""Volume-preserving transformer :math:`y_i = x_i + b_i` (``log_det_J = 0``).

    Features listed in ``periodic_indices`` are wrapped as ``(x + b) % (limits[1] - limits[0]) + limits[0]``
    (exactly the reference's expression, affine.py:415-416).
    """

    n_parameters_per_feature = 1

    def __init__(self, periodic_indices=None, periodic_limits=None):
        super().__init__()
        self.periodic_indices = periodic_indices
        self.periodic_limits = periodic_limits
        self._specs = {}

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        return torch.zeros(size=(self.n_parameters_per_feature * n_features,))

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.tile((self.n_parameters_per_feature,))

    def _parts(self, n_features):
        if n_features not in self._specs:
            self._specs[n_features] = _ShiftSpec(n_features, self.periodic_indices, self.periodic_limits)
        return [_program.Part('shift', self._specs[n_features], n_features, 1)]


def volume_preserving_shift_transformer(x, shift, periodic_indices=None, periodic_limits=None):
    """Functional form (reference affine.py:366-418): returns ``(y, log_det_J)`` with ``log_det_J = 0``."""
    t = VolumePreservingShiftTransformer(periodic_indices, periodic_limits)
    return _program.run(t._parts(x.shape[1]), x, shift)


def volume_preserving_shift_transformer_inverse(y, shift, periodic_indices=None, periodic_limits=None):
    """Inverse of :func:`volume_preserving_shift_transformer` (reference affine.py:421-456)."""
    t = VolumePreservingShiftTransformer(periodic_indices, periodic_limits)
    return _program.run(t._parts(y.shape[1]), y, shift, inverse=True)


def affine_transformer(x, shift, log_scale):
    """Functional form (reference affine.py:281-323): returns ``(y, log_det_J)``."""
    par = torch.cat([shift, log_scale], dim=1)
    return _program.run(AffineTransformer()._parts(x.shape[1]), x, par)


def affine_transformer_inverse(y, shift, log_scale):
    """Inverse of :func:`affine_transformer` (reference affine.py:326-363)."""
    par = torch.cat([shift, log_scale], dim=1)
    return _program.run(AffineTransformer()._parts(y.shape[1]), y, par, inverse=True)
