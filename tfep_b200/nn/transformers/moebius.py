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
        # unit_sphere doubles as the kernel variant: 0 radius |x|, 1 unit sphere, 2 symmetrized
        self.dimension, self.max_radius, self.unit_sphere = int(dimension), float(max_radius), int(unit_sphere)


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


class SymmetrizedMoebiusTransformer(MAFTransformer):
    r"""Symmetrized Moebius transformation (reference tfep/nn/transformers/moebius.py:193-372, 481-629)

    :math:`y = |x| \frac{f(x; w) + f(x; -w)}{|f(x; w) + f(x; -w)|}` with :math:`f` the Moebius map on the sphere of
    radius ``|x|``; closed-form ``log_det_J`` and analytic inverse.  Same kernels as :class:`MoebiusTransformer`
    (tfepb_moebius / tfepb_moebius_backward, variant 2).

    ``get_identity_parameters`` returns a tiny random tensor (``identity_eps``) rather than zeros, because the
    gradient w.r.t. the parameters vanishes at exactly zero.
    """

    def __init__(self, dimension: int, max_radius: float = 0.99, identity_eps: float = 1e-9):
        super().__init__()
        self.dimension = dimension
        self.max_radius = max_radius
        self.identity_eps = identity_eps

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        par = torch.rand(n_features)
        return (2 * par - 1) * self.identity_eps

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.detach().clone()

    def _parts(self, n_features):
        spec = _MoebiusSpec(self.dimension, self.max_radius, 2)
        return [_program.Part('moebius', spec, n_features, 1, param_major=False)]


def symmetrized_moebius_transformer(x, w, max_radius=0.99):
    """Functional form on ``(batch, n_vectors, dimension)`` tensors (reference moebius.py:481-550)."""
    batch, n_vectors, dimension = x.shape
    t = SymmetrizedMoebiusTransformer(dimension, max_radius)
    y, ld = _program.run(t._parts(n_vectors * dimension), x.reshape(batch, -1), w.reshape(batch, -1))
    return y.reshape(batch, n_vectors, dimension), ld


def symmetrized_moebius_transformer_inverse(x, w, max_radius=0.99):
    """Inverse of :func:`symmetrized_moebius_transformer` (reference moebius.py:553-600)."""
    batch, n_vectors, dimension = x.shape
    t = SymmetrizedMoebiusTransformer(dimension, max_radius)
    y, ld = _program.run(t._parts(n_vectors * dimension), x.reshape(batch, -1), w.reshape(batch, -1), inverse=True)
    return y.reshape(batch, n_vectors, dimension), ld
