"""Moebius transformer on the sphere of radius ``|x|`` (reference tfep/nn/transformers/moebius.py:27-190, 374-478).

Kernels: tfepb_moebius / tfepb_moebius_backward.  The log-det uses the closed form
``(d - 1) log((|x|^2 - |w|^2) / |x - w|^2)`` per vector (``d log(.)`` with ``unit_sphere=True``), which equals
the ``slogdet`` of the explicit Jacobian the reference builds (moebius.py:463-476); see DESIGN.md.
"""

import torch

from ... import _program
from .transformer import MAFTransformer


class _MoebiusSpec:
    def __init__(self, dimension, max_radius, unit_sphere):
        self.dimension, self.max_radius, self.unit_sphere = int(dimension), float(max_radius), bool(unit_sphere)


class MoebiusTransformer(MAFTransformer):
    r"""Moebius transformation :math:`y = \frac{|x|^2 - |w|^2}{|x - w|^2} (x - w) - w` of ``dimension``-dimensional
    vector blocks (consecutive features).  The parameters ``w`` are rescaled so that ``|w| < max_radius |x|``.
    """

    def __init__(self, dimension: int, max_radius: float = 0.99, unit_sphere: bool = False):
        super().__init__()
        self.dimension = dimension
        self.max_radius = max_radius
        self.unit_sphere = unit_sphere

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        return torch.zeros(size=(n_features,))

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.detach().clone()

    def _parts(self, n_features):
        spec = _MoebiusSpec(self.dimension, self.max_radius, self.unit_sphere)
        return [_program.Part('moebius', spec, n_features, 1, param_major=False)]


def moebius_transformer(x, w, max_radius=0.99, unit_sphere=False, return_log_det_J=True):
    """Functional form on ``(batch, n_vectors, dimension)`` tensors (reference moebius.py:374-478)."""
    batch, n_vectors, dimension = x.shape
    t = MoebiusTransformer(dimension, max_radius, unit_sphere)
    y, ld = _program.run(t._parts(n_vectors * dimension), x.reshape(batch, -1), w.reshape(batch, -1))
    y = y.reshape(batch, n_vectors, dimension)
    return (y, ld) if return_log_det_J else y
