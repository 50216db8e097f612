"""Interface of a conditioner usable by :class:`tfep_b200.nn.flows.AutoregressiveFlow`.

Reference: tfep/nn/conditioners/conditioner.py:26-63.
"""

import abc

import torch


class Conditioner(abc.ABC, torch.nn.Module):
    """A conditioner maps the input features to the parameters of the transformer."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``(batch, n_features) -> (batch, n_parameters)``."""
        return super().forward(x)

    @abc.abstractmethod
    def set_output(self, output: torch.Tensor):
        """Initialise the conditioner so that it returns ``output`` (shape ``(n_parameters,)``) for every input."""
