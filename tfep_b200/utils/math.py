"""Batch-wise products and an autograd Jacobian used by the tests (reference: tfep/utils/math.py)."""

import torch


def batchwise_dot(x1, x2, keepdim=False):
    """Dot product over the last dimension (reference tfep/utils/math.py:24-44)."""
    return (x1 * x2).sum(dim=-1, keepdim=keepdim)


def batchwise_outer(x1, x2):
    """Outer product over the last dimension (reference tfep/utils/math.py:47-66)."""
    return torch.einsum('...i,...j->...ij', x1, x2)


def batch_autograd_log_abs_det_J(x, y):
    """log|det dy/dx| per sample from autograd, one output column at a time.

    Reference: tfep/utils/math.py:178-203.  Test helper: O(D) backward passes.
    """
    batch, n = y.shape[0], y.shape[1]
    jac = torch.empty(batch, n, x.shape[1], dtype=x.dtype, device=x.device)
    for i in range(n):
        g = torch.zeros_like(y)
        g[:, i] = 1.0
        jac[:, i] = torch.autograd.grad(y, x, g, retain_graph=True)[0]
    return torch.linalg.slogdet(jac)[1]
