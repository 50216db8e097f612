"""CPU restatement of the reference (T)FEP estimator and bootstrap.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinned: yes --
against the real reference (oracle/check_against_reference.py) and the golden
index streams / statistics in tests/golden/analysis.npz.

The resample indices of the reference come from ``torch.randint`` on a CPU
``torch.Generator`` (analysis/bootstrap.py:214-218).  PyTorch's CPU generator
is the 32-bit Mersenne Twister MT19937 seeded with ``init_genrand(seed)``
(published algorithm: Matsumoto & Nishimura 1998, ACM TOMACS 8(1); ATen
``at::mt19937``), and for ranges below 2**32 ``randint`` returns
``u32 % range`` for consecutive outputs, row-major over the requested shape.
``Mt19937`` below restates that generator with numpy so that the index stream
can be reproduced without torch.
"""

import numpy as np
import torch

_N, _M = 624, 397
_UPPER, _LOWER, _MATRIX_A = np.uint32(0x80000000), np.uint32(0x7fffffff), np.uint32(0x9908b0df)


class Mt19937:
    """MT19937 raw 32-bit stream, identical to ``torch.Generator().manual_seed(seed)`` on CPU."""

    def __init__(self, seed: int):
        s = np.empty(_N, dtype=np.uint64)
        s[0] = seed & 0xffffffff
        for i in range(1, _N):          # init_genrand
            s[i] = (1812433253 * (s[i - 1] ^ (s[i - 1] >> np.uint64(30))) + i) & 0xffffffff
        self.state = s.astype(np.uint32)
        self.pos = _N

    @staticmethod
    def next_state(st: np.ndarray) -> np.ndarray:
        """One full twist: the next 624 state words from the current 624."""
        new = st.copy()

        def mix(a, b):
            yv = (a & _UPPER) | (b & _LOWER)
            return (yv >> np.uint32(1)) ^ np.where(yv & np.uint32(1), _MATRIX_A, np.uint32(0))

        # new[i] = new_or_old[i + M] ^ mix(old[i], old[i+1]); three dependency-free spans.
        new[0:_N - _M] = st[_M:_N] ^ mix(st[0:_N - _M], st[1:_N - _M + 1])
        for lo in range(_N - _M, _N - 1, _N - _M):
            hi = min(lo + (_N - _M), _N - 1)
            new[lo:hi] = new[lo - (_N - _M):hi - (_N - _M)] ^ mix(st[lo:hi], st[lo + 1:hi + 1])
        new[_N - 1] = new[_M - 1] ^ mix(st[_N - 1:_N], new[0:1])[0]
        return new

    @staticmethod
    def temper(y: np.ndarray) -> np.ndarray:
        y = y ^ (y >> np.uint32(11))
        y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
        y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000))
        return y ^ (y >> np.uint32(18))

    def raw(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        k = 0
        while k < n:
            if self.pos == _N:
                self.state = self.next_state(self.state)
                self.pos = 0
            take = min(n - k, _N - self.pos)
            out[k:k + take] = self.temper(self.state[self.pos:self.pos + take])
            self.pos += take
            k += take
        return out


def resample_indices(seed: int, n_resamples: int, sample_size: int, max_idx: int, skip: int = 0) -> np.ndarray:
    """Index matrix ``(n_resamples, sample_size)`` the reference draws.  analysis/bootstrap.py:200-218.

    Consecutive ``batch`` chunks of the reference continue the same stream, so
    the matrix does not depend on ``batch``.  ``skip`` raw outputs are discarded
    first (a generator that was already used).
    """
    assert max_idx < 2**32
    g = Mt19937(seed)
    if skip:
        g.raw(skip)
    return (g.raw(n_resamples * sample_size).astype(np.int64) % max_idx).reshape(n_resamples, sample_size)


def fep_estimator(data, kT=1.0, weights=None, vectorized=False):
    """analysis/estimator.py:24-86.  NB: log(n) is evaluated in the default dtype from an
    int64 tensor (:75-77) and biased data is laid out (n_samples, 2) (:66,71)."""
    if vectorized:
        work, bias = (data, None) if data.dim() == 2 else data.permute(2, 0, 1)
    else:
        work, bias = (data, None) if data.dim() == 1 else data.T
    if bias is None:
        if weights is None:
            log_w = -torch.log(torch.tensor(work.shape[-1]))
        else:
            log_w = torch.log(weights)
    elif weights is not None:
        raise NotImplementedError('Bayesian bootstrapping is not supported with biased data.')
    else:
        log_w = torch.nn.functional.log_softmax(bias / kT, dim=-1)
    return -kT * torch.logsumexp(-work / kT + log_w, dim=-1)


def bootstrap_statistics(data, statistic, n_resamples, sample_size=None, take_first_only=False,
                         batch=None, generator=None):
    """The per-resample statistics.  analysis/bootstrap.py:185-233."""
    n = len(data)
    sample_size = n if sample_size is None else sample_size
    batch = n_resamples if batch is None else batch
    max_idx = sample_size if take_first_only else n
    out = torch.empty(n_resamples, dtype=data.dtype)
    for k in range(0, n_resamples, batch):
        nb = min(batch, n_resamples - k)
        idx = torch.randint(low=0, high=max_idx, size=(nb, sample_size), generator=generator)
        expanded = data.expand((nb, *data.shape))
        if data.dim() > 1:
            idx = idx.repeat_interleave(repeats=data.shape[1], dim=1).reshape(nb, sample_size, data.shape[1])
        out[k:k + nb] = statistic(torch.gather(expanded, dim=1, index=idx), vectorized=True)
    return out


def summarize(stats, confidence_level=0.95, method='percentile', full_statistic=None):
    """analysis/bootstrap.py:163-178."""
    alpha = (1 - confidence_level) / 2
    lo, hi = torch.quantile(stats, q=torch.tensor([alpha, 1 - alpha], dtype=stats.dtype))
    if method == 'basic':
        lo, hi = 2 * full_statistic - hi, 2 * full_statistic - lo
    return dict(confidence_interval=dict(low=lo, high=hi), standard_deviation=torch.std(stats),
                mean=torch.mean(stats), median=torch.median(stats))


def bayesian_bootstrap_statistics(data, statistic, n_resamples, sample_size=None, batch=None):
    """Statistics under Dirichlet(1, ..., 1) sample weights.  analysis/bootstrap.py:236-262 (the weights come
    from the GLOBAL torch generator: seed with torch.manual_seed for reproducibility)."""
    n = len(data) if sample_size is None else sample_size
    batch = n_resamples if batch is None else batch
    out = torch.empty(n_resamples, dtype=data.dtype)
    dirichlet = torch.distributions.Dirichlet(torch.ones(n))
    for k in range(0, n_resamples, batch):
        nb = min(batch, n_resamples - k)
        weights = dirichlet.sample((nb,))
        out[k:k + nb] = statistic(data.expand((nb, *data.shape))[:, :n], weights=weights, vectorized=True)
    return out


def bootstrap(data, statistic, *, confidence_level=0.95, n_resamples=9999, bootstrap_sample_size=None,
              take_first_only=False, batch=None, method='percentile', bayesian=False, generator=None):
    """analysis/bootstrap.py:24-182."""
    if bayesian and generator is not None:
        raise ValueError('Bayesian bootstrapping does not support random number generators.')
    if bayesian and bootstrap_sample_size is not None and not take_first_only:
        raise ValueError('With Bayesian bootstrapping, specifying a bootstrap_sample_size '
                         'is supported only when take_first_only is True.')
    sizes = [len(data)] if bootstrap_sample_size is None else list(bootstrap_sample_size)
    res = []
    with torch.no_grad():
        for s in sizes:
            if bayesian:
                st = bayesian_bootstrap_statistics(data, statistic, n_resamples, s, batch)
            else:
                st = bootstrap_statistics(data, statistic, n_resamples, s, take_first_only, batch, generator)
            full = statistic(data.unsqueeze(0)) if method == 'basic' else None
            res.append(summarize(st, confidence_level, method, full))
    return res[0] if len(sizes) == 1 else res
