"""MT19937 jump-ahead, host part (tfepb_mt19937_jump_polynomial): t^J mod phi, with phi recovered by Berlekamp-Massey
in the library, applied to a generator state with a numpy restatement of the kernel's block evaluation, must land on the
window the sequential generator reaches after J steps -- all 624 x 32 bits (reference stream: bootstrap.py:207-218 =
torch.randint on a CPU generator = raw MT19937 words)."""

import numpy as np
import pytest

from oracle.analysis_oracle import Mt19937
from tfep_b200 import _ops

N, M = 624, 397


def _raw_words(seed, n):
    """Untempered words x_0 .. x_{n-1} of the stream (x_0 = first word output after seeding)."""
    g = Mt19937(seed)
    out = np.empty(n, dtype=np.uint32)
    k = 0
    while k < n:
        g.state = g.next_state(g.state)
        take = min(N, n - k)
        out[k:k + take] = g.state[:take]
        k += take
    return out


def _apply(poly_words, window):
    """F g(F) window with g given as 19968 bits: Horner over 32 blocks of 624 coefficients, F^624 = one twist,
    sum_j g_j F^j W = XOR of shifted copies of (W, twist(W)); then one real step (exact on the low bits of word 0)."""
    bits = np.unpackbits(poly_words.view(np.uint8), bitorder='little')
    run = np.concatenate([window, Mt19937.next_state(window)])
    h = np.zeros(N, dtype=np.uint32)
    for c in range(31, -1, -1):
        h = Mt19937.next_state(h)
        for j in np.nonzero(bits[c * N:(c + 1) * N])[0]:
            h ^= run[j:j + N]
    y = (h[0] & np.uint32(0x80000000)) | (h[1] & np.uint32(0x7fffffff))
    new = h[M] ^ (y >> np.uint32(1)) ^ (np.uint32(0x9908b0df) if y & np.uint32(1) else np.uint32(0))
    return np.concatenate([h[1:], [new]]).astype(np.uint32)


@pytest.mark.parametrize('seed,J', [(1, 1), (5489, 2), (7, 623), (7, 624), (99, 625), (12345, 100003), (3, 1 << 20)])
def test_jump_polynomial_lands_on_the_sequential_window(seed, J):
    words = _raw_words(seed, J + N)
    poly = _ops.mt19937_jump_polynomial(J - 1)            # t^(J-1): the kernel takes the last step for real
    assert np.array_equal(_apply(poly, words[:N]), words[J:J + N])


def test_polynomial_algebra():
    """t^a t^b = t^(a+b) mod phi, checked through the action on a state; degree of every residue < 19937."""
    a, b = 700001, 1234567
    w = _raw_words(11, N)
    wa = _apply(_ops.mt19937_jump_polynomial(a - 1), w)
    wab = _apply(_ops.mt19937_jump_polynomial(b - 1), wa)
    assert np.array_equal(wab, _apply(_ops.mt19937_jump_polynomial(a + b - 1), w))
    bits = np.unpackbits(_ops.mt19937_jump_polynomial(10 ** 11).view(np.uint8), bitorder='little')
    assert not bits[19937:].any() and bits.sum() > 9000     # a dense residue
