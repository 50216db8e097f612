"""Chain of normalizing flows (reference tfep/nn/flows/sequential.py:24-68)."""

import torch


class SequentialFlow(torch.nn.Sequential):
    """Runs the wrapped flows in order (reverse order for ``inverse``) and sums their log-det Jacobians."""

    def n_parameters(self):
        """The total number of parameters that can be optimized."""
        return sum(flow.n_parameters() for flow in self)

    def forward(self, x):
        return self._pass(x, inverse=False)

    def inverse(self, y):
        return self._pass(y, inverse=True)

    def _pass(self, x, inverse):
        cumulative_log_det_J = None
        for flow in (reversed(self) if inverse else self):
            x, log_det_J = flow.inverse(x) if inverse else flow(x)
            cumulative_log_det_J = log_det_J if cumulative_log_det_J is None else cumulative_log_det_J + log_det_J
        if cumulative_log_det_J is None:
            cumulative_log_det_J = torch.zeros(x.size(0), dtype=x.dtype, device=x.device)
        return x, cumulative_log_det_J
