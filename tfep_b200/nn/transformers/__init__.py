"""Transformers for autoregressive normalizing flows (reference tfep/nn/transformers/__init__.py)."""

from .affine import (AffineTransformer, VolumePreservingShiftTransformer, affine_transformer, affine_transformer_inverse,
                     volume_preserving_shift_transformer, volume_preserving_shift_transformer_inverse)
from .mixed import MixedTransformer
from .moebius import (MoebiusTransformer, SymmetrizedMoebiusTransformer, moebius_transformer,
                      symmetrized_moebius_transformer, symmetrized_moebius_transformer_inverse)
from .sos import SOSPolynomialTransformer, SOSPolynomialTransformerFunc, sos_polynomial_transformer
from .spline import NeuralSplineTransformer, neural_spline_transformer
from .transformer import MAFTransformer, Transformer
