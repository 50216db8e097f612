"""Flow wrapper working in PCA-whitened coordinates (reference tfep/nn/flows/pca.py:25-125).

The whitening / blackening maps are two dense (batch, D) x (D, D) products either side of the wrapped flow; they run on
the package's own exact GEMM kernels (tfepb_masked_linear_forward / _backward_input through ``masked_linear`` without a
mask), differentiable like every other layer.  The matrices are estimated once, at construction, from data.
"""

import torch

from ..masked import masked_linear


def _cov(x):
    """Sample covariance (ddof = 1) and mean of the rows of ``x`` (reference utils/math.py:67-134)."""
    mean = torch.mean(x, 0)
    xc = x - mean
    return torch.matmul(xc.t(), xc) / (x.shape[0] - 1), mean


class PCAWhitenedFlow(torch.nn.Module):
    """Pass PCA-whitened coordinates to the wrapped flow and (optionally) blacken its output.

    Parameters
    ----------
    flow : torch.nn.Module
    x : torch.Tensor
        ``(n_samples, n_features)`` data from which mean and covariance are estimated.
    blacken : bool
        If False the output stays in whitened space and the log-det of the whitening map is added (subtracted in the
        inverse direction).
    """

    def __init__(self, flow, x, blacken=True):
        super().__init__()
        self.flow = flow
        self.blacken = blacken
        cov, mean = _cov(x.detach())
        eigvalues, eigvectors = torch.linalg.eigh(cov)
        if torch.any(eigvalues < 0.0):
            raise ValueError(
                'Cannot determine the PCA whitening matrix since some of the '
                'eigenvalues of the covariance matrix estimate are negative. '
                'Likely, this is due to an insufficient number of samples.')
        singular_values = torch.sqrt(eigvalues)
        self.register_buffer('mean', mean)
        self.register_buffer('whitening_matrix', torch.matmul(eigvectors, torch.diag(1. / singular_values)))
        self.register_buffer('blackening_matrix', torch.matmul(torch.diag(singular_values), eigvectors.t()))
        self.register_buffer('whitening_log_det_J', -torch.sum(torch.log(singular_values)))

    def n_parameters(self):
        return self.flow.n_parameters()

    def forward(self, x):
        return self._pass(x, inverse=False)

    def inverse(self, y):
        return self._pass(y, inverse=True)

    def _whiten(self, x):
        # (x - mean) W = x W - mean W: one GEMM with the bias folded in
        w_t = self.whitening_matrix.t().contiguous()
        return masked_linear(x, w_t, -torch.mv(w_t, self.mean))

    def _blacken(self, x):
        return masked_linear(x, self.blackening_matrix.t().contiguous(), self.mean)

    def _pass(self, x, inverse):
        whiten = not inverse or self.blacken
        blacken = inverse or self.blacken
        if whiten:
            x = self._whiten(x)
        y, log_det_J = self.flow.inverse(x) if inverse else self.flow(x)
        if blacken:
            y = self._blacken(y)
        if not (whiten and blacken):
            log_det_J = log_det_J + self.whitening_log_det_J if whiten else log_det_J - self.whitening_log_det_J
        return y, log_det_J
