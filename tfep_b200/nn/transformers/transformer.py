"""Transformer interfaces (reference tfep/nn/transformers/transformer.py:26-127).

A transformer maps features ``x`` to ``y`` given per-sample parameters produced by a conditioner and
returns ``log|det dy/dx|``.  On top of the reference's abstract API every transformer here can describe
itself as a kernel program (:meth:`Transformer._parts`), which is what the flows execute.
"""

import abc

import torch

from ... import _program


class Transformer(abc.ABC, torch.nn.Module):
    """A transformer for an autoregressive flow."""

    def forward(self, x: torch.Tensor, parameters: torch.Tensor):
        """``(batch, n_features), (batch, n_parameters) -> (y, log_det_J)``."""
        return _program.run(self._parts(x.shape[1]), x, parameters, inverse=False)

    def inverse(self, y: torch.Tensor, parameters: torch.Tensor):
        """Inverse map; returns ``(x, log_det_J)`` of the inverse."""
        return _program.run(self._parts(y.shape[1]), y, parameters, inverse=True)

    @abc.abstractmethod
    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        """Parameters (shape ``(n_parameters,)``) that make the transformer the identity."""

    @abc.abstractmethod
    def _parts(self, n_features: int):
        """List of :class:`tfep_b200._program.Part` lowering this transformer for ``n_features`` inputs."""


class MAFTransformer(Transformer):
    """A transformer usable inside :class:`tfep_b200.nn.flows.MAF`."""

    @abc.abstractmethod
    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        """Degrees of the conditioner outputs feeding this transformer, given the degrees of its inputs."""
