"""TEST-ONLY host build of the transformer math headers (tfep_b200/csrc/tx_math.cuh).

Compiles tests/hostcheck/tx_hostcheck.cpp with g++ on first use and exposes thin ctypes wrappers
working on CPU torch tensors in the reference parameter layout (B, P, F).  This lets the CPU test
suite check the exact formulas the CUDA kernels execute against the oracle without a GPU.  It is
not part of the product: tfep_b200 never loads it.
"""

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'tx_hostcheck.cpp')
_HDRS = [os.path.join(_HERE, '..', '..', 'tfep_b200', 'csrc', h) for h in ('tx_math.cuh', 'hd_math.cuh')]
_OUT = os.path.join(_HERE, '_build', 'libtxhost.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(_OUT)) or any(os.path.getmtime(p) > os.path.getmtime(_OUT) for p in [_SRC] + _HDRS)
        if stale:
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            subprocess.check_call(['g++', '-O2', '-x', 'c++', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                                   '-o', _OUT, _SRC])
        _lib = ctypes.CDLL(_OUT)
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _io(x, par, gy=None, gl=None):
    B, F = x.shape
    P = par.shape[1] // F
    x, par = x.contiguous(), par.contiguous()
    y, ld = torch.empty_like(x), torch.empty(B, dtype=x.dtype)
    gx, gpar = torch.empty_like(x), torch.zeros_like(par)
    if gy is not None:
        gy = gy.contiguous()
        gl = (torch.zeros(B, dtype=x.dtype) if gl is None else gl).contiguous()
    keep = (x, par, y, ld, gy, gl, gx, gpar)
    args = [ctypes.c_int(int(x.dtype == torch.float64)), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(P),
            _p(x), _p(par), _p(y), _p(ld), _p(gy), _p(gl), _p(gx), _p(gpar)]
    return args, keep


def affine(x, par, inverse=False, gy=None, gl=None):
    args, k = _io(x, par, gy, gl)
    lib().hc_affine(*args, ctypes.c_int(int(inverse)), ctypes.c_int(int(gy is not None)))
    return (k[6], k[7]) if gy is not None else (k[2], k[3])


def shift(x, par, spec, inverse=False, gy=None, gl=None):
    """``spec``: oracle.flow_oracle.Shift."""
    args, k = _io(x, par, gy, gl)
    F = x.shape[1]
    period, lower = torch.zeros(F, dtype=x.dtype), torch.zeros(F, dtype=x.dtype)
    if spec.periodic_indices is not None:
        lim = spec.periodic_limits.to(x.dtype)
        period[spec.periodic_indices] = lim[1] - lim[0]
        lower[spec.periodic_indices] = lim[0]
    lib().hc_shift(*args, _p(period), _p(lower), ctypes.c_int(int(inverse)), ctypes.c_int(int(gy is not None)))
    return (k[6], k[7]) if gy is not None else (k[2], k[3])


def sos(x, par, n_poly, gy=None):
    args, k = _io(x, par, gy, None)
    lib().hc_sos(*args, ctypes.c_int(n_poly), ctypes.c_int(int(gy is not None)))
    return (k[6], k[7]) if gy is not None else (k[2], k[3])


def moebius(x, par, d, max_radius=0.99, unit_sphere=False, inverse=False, gy=None, gl=None):
    args, k = _io(x, par, gy, gl)
    lib().hc_moebius(*args, ctypes.c_int(d), ctypes.c_double(max_radius), ctypes.c_int(int(unit_sphere)),
                     ctypes.c_int(int(inverse)), ctypes.c_int(int(gy is not None)))
    return (k[6], k[7]) if gy is not None else (k[2], k[3])


def spline(x, par, spec, inverse=False, gy=None, gl=None, return_bins=False):
    """``spec`` is any object with the attributes of oracle.flow_oracle.Spline."""
    args, k = _io(x, par, gy, gl)
    dt = x.dtype
    dom = [t.to(dt).contiguous() for t in (spec.x0, spec.xf, spec.y0, spec.yf)]
    bins = torch.empty(x.shape, dtype=torch.int32) if return_bins else None
    lib().hc_spline(*args, ctypes.c_int(spec.n_bins), ctypes.c_int(int(spec.circular)),
                    ctypes.c_int(int(spec.identity_boundary_slopes)), ctypes.c_int(int(spec.learn_lower_bound)),
                    ctypes.c_int(int(spec.learn_upper_bound)), *[_p(t) for t in dom],
                    ctypes.c_double(spec.min_bin_size), ctypes.c_double(spec.min_slope),
                    ctypes.c_int(int(inverse)), ctypes.c_int(int(gy is not None)), _p(bins))
    if gy is not None:
        return k[6], k[7]
    return (k[2], k[3], bins) if return_bins else (k[2], k[3])
