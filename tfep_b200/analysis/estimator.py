"""(Targeted) free energy perturbation estimator (reference tfep/analysis/estimator.py:24-86).

``Delta f = -kT logsumexp(-w / kT + log_weights)`` evaluated by the single-pass online log-sum-exp kernel
(tfepb_lse): one read of the work values, (max, sum) carried in double.  At N > 1 ranks each rank reduces
its shard and the (max, sum) pairs are combined with :func:`combine_partials` (one tiny all-gather).
"""

import math

import torch

from .. import _ops


def _log_n(n):
    # The reference evaluates log(n) in the DEFAULT dtype from an integer tensor (estimator.py:75-77),
    # i.e. in float32 unless the default is changed, even for float64 data.  Reproduced on purpose.
    return float(torch.log(torch.tensor(n)))


def combine_partials(partials):
    """Combine ``(..., 2)`` pairs ``(max_i, sum_i)`` of shards of one data set: returns ``(max, sum)`` with
    ``sum = sum_i sum_i * exp(max_i - max)``.  Works on CPU or CUDA tensors (host logic of the N-rank path)."""
    m_i, s_i = partials[..., 0], partials[..., 1]
    m = m_i.max(dim=-1).values
    s = (s_i * torch.exp(m_i - m.unsqueeze(-1))).sum(dim=-1)
    return m, s


def lse_partial(work, kT=1.0, log_weights=None):
    """``(max, sum exp(v - max))`` of ``v = -work / kT (+ log_weights)`` as a (2,) float64 device tensor."""
    return _ops.lse(work, -1.0 / kT, log_weights)


def _row_estimate(work, kT, log_weights, log_norm):
    if not torch.is_tensor(log_norm) and work.dtype in (torch.float32, torch.float64):
        # the usual case: the final reduction of the kernel evaluates the estimate itself (two launches in all)
        return _ops.fep_estimate(work, kT, log_norm, log_weights)[0]
    o = lse_partial(work, kT, log_weights)
    return (-kT * (o[0] + torch.log(o[1]) - log_norm)).to(work.dtype)


def fep_estimator(data, kT=1.0, weights=None, vectorized=False):
    """FEP estimator.

    Parameters
    ----------
    data : torch.Tensor
        Work values ``(n_samples,)`` in units of kT, or ``(n_samples, 2)`` = (work, bias) for reweighting
        (the layout the reference actually consumes, estimator.py:66,71).  With ``vectorized`` a leading
        bootstrap dimension is expected.
    kT : float
    weights : torch.Tensor, optional
        ``(n_bootstraps, n_samples)`` sample weights for Bayesian bootstrapping.
    vectorized : bool

    Returns
    -------
    df : torch.Tensor
        0-dim (or ``(n_bootstraps,)`` if ``vectorized``) tensor on ``data.device``.
    """
    if vectorized:
        work, bias = (data, None) if data.dim() == 2 else (data[..., 0], data[..., 1])
        rows = work.shape[0]
    else:
        work, bias = (data, None) if data.dim() == 1 else (data[:, 0], data[:, 1])
        rows = None
    if bias is not None and weights is not None:
        raise NotImplementedError('Bayesian bootstrapping is not supported with biased data.')

    def one(w, b, wts):
        if b is None:
            if wts is None:
                return _row_estimate(w, kT, None, _log_n(w.shape[-1]))
            return _row_estimate(w, kT, torch.log(wts), 0.0)
        # log_softmax(bias / kT) = bias / kT - logsumexp(bias / kT)
        ob = _ops.lse(b, 1.0 / kT)
        return _row_estimate(w, kT, b / kT, ob[0] + torch.log(ob[1]))

    if rows is None:
        return one(work, bias, weights)
    if rows > 4:
        # many rows (the generic-statistic route of bootstrap: `statistic(samples, vectorized=True)`): one batched
        # reduction over the (rows, n) matrix instead of two kernel launches per row
        v = -work / kT
        if bias is not None:
            v = v + torch.log_softmax(bias / kT, dim=-1)
            return (-kT * torch.logsumexp(v, dim=-1)).to(work.dtype)
        if weights is not None:
            return (-kT * torch.logsumexp(v + torch.log(weights), dim=-1)).to(work.dtype)
        return (-kT * (torch.logsumexp(v, dim=-1) - _log_n(work.shape[-1]))).to(work.dtype)
    return torch.stack([one(work[r], None if bias is None else bias[r], None if weights is None else weights[r])
                        for r in range(rows)])
