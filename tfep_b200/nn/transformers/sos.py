"""Sum-of-squares polynomial transformer (reference tfep/nn/transformers/sos.py:28-306).

``y = a0 + int_0^x sum_k (a_k0 + a_k1 z)^2 dz``.  Kernels: tfepb_sos / tfepb_sos_backward.  As in the
reference there is no inverse (sos.py:111-114) and the log-det output carries no gradient (sos.py:233).
"""

import numpy as np
import torch

from ... import _program
from .transformer import MAFTransformer


class _SosSpec:
    def __init__(self, n_polynomials):
        self.n_polynomials = int(n_polynomials)


class SOSPolynomialTransformer(MAFTransformer):
    """Sum-of-squares polynomial transformer with ``n_polynomials`` squared first-degree polynomials.

    ``parameters`` reshapes to ``(batch, 1 + 2 K, n_features)``: ``a0`` followed by ``a_10, a_11, ..., a_K0, a_K1``.
    """

    def __init__(self, n_polynomials=2):
        super().__init__()
        if n_polynomials < 2:
            raise ValueError('n_polynomials must be strictly greater than 1.')
        self.n_polynomials = n_polynomials

    @property
    def degree_polynomials(self):
        return 1

    @property
    def parameters_per_polynomial(self):
        return self.degree_polynomials + 1

    @property
    def n_parameters_per_feature(self):
        return self.parameters_per_polynomial * self.n_polynomials + 1

    def inverse(self, y, parameters):
        raise NotImplementedError('Inversion of SOS polynomial transformer has not been implemented yet.')

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        par = torch.zeros(size=(self.n_parameters_per_feature, n_features))
        par[1::self.parameters_per_polynomial].fill_(np.sqrt(1 / self.n_polynomials))
        return par.flatten()

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.tile((self.n_parameters_per_feature,))

    def _parts(self, n_features):
        return [_program.Part('sos', _SosSpec(self.n_polynomials), n_features, self.n_parameters_per_feature)]


def sos_polynomial_transformer(x, parameters):
    """Functional form: ``parameters`` has shape ``(batch, 1 + 2 K, n_features)`` (reference sos.py:163-306)."""
    batch, n_par, n_features = parameters.shape
    t = SOSPolynomialTransformer((n_par - 1) // 2)
    return _program.run(t._parts(n_features), x, parameters.reshape(batch, -1))


class SOSPolynomialTransformerFunc:
    """Name kept for callers of the reference's ``SOSPolynomialTransformerFunc.apply(x, parameters)`` (sos.py:207-310):
    the forward and the hand-written backward live in the CUDA kernels (tfepb_sos / tfepb_sos_backward)."""

    apply = staticmethod(sos_polynomial_transformer)

    @staticmethod
    def get_sos_poly_coefficients(parameters):
        """List of the four ``(batch, n_features)`` coefficients ``a0, c1, c2, c3`` of ``y = a0 + c1 x + c2 x^2 + c3 x^3``
        from the ``(batch, 1 + 2 K, n_features)`` parameters (reference sos.py:271-306; host-side helper, tensor algebra)."""
        k0, k1 = parameters[:, 1::2], parameters[:, 2::2]
        return [parameters[:, 0], (k0 * k0).sum(dim=1), (k0 * k1).sum(dim=1), (k1 * k1).sum(dim=1) / 3]
