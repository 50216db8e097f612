"""KL-divergence training loss between Boltzmann distributions (reference tfep/loss.py:26-140)."""

from typing import Optional

import torch


class BoltzmannKLDivLoss(torch.nn.Module):
    """``mean_i [u_B(x_i) - log|det J(x_i)| - u_A(x_i)]``, optionally weighted by ``softmax(log_weights)``.

    A reduction over ``(batch,)`` vectors: plain PyTorch (plumbing around the flow kernels).
    """

    def __init__(self, ignore_nan: bool = False):
        super().__init__()
        self.ignore_nan = ignore_nan

    def forward(self, target_potentials: torch.Tensor, log_det_J: Optional[torch.Tensor] = None,
                log_weights: Optional[torch.Tensor] = None, ref_potentials: Optional[torch.Tensor] = None):
        reduced_work = target_potentials
        if log_det_J is not None:
            reduced_work = reduced_work - log_det_J
        if ref_potentials is not None:
            reduced_work = reduced_work - ref_potentials
        if log_weights is not None:
            weights = torch.nn.functional.softmax(log_weights, dim=0)
            if self.ignore_nan:
                return torch.nansum(weights * reduced_work)
            return torch.sum(weights * reduced_work)
        if self.ignore_nan:
            return torch.nanmean(reduced_work)
        return torch.mean(reduced_work)
