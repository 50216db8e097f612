"""Affine transformer ``y = x exp(a) + b`` (reference tfep/nn/transformers/affine.py:28-141, 281-363).

Kernel: tfepb_affine / tfepb_affine_backward.
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


def affine_transformer(x, shift, log_scale):
    """Functional form (reference affine.py:281-323): returns ``(y, log_det_J)``."""
    par = torch.cat([shift, log_scale], dim=1)
    return _program.run(AffineTransformer()._parts(x.shape[1]), x, par)


def affine_transformer_inverse(y, shift, log_scale):
    """Inverse of :func:`affine_transformer` (reference affine.py:326-363)."""
    par = torch.cat([shift, log_scale], dim=1)
    return _program.run(AffineTransformer()._parts(y.shape[1]), y, par, inverse=True)
