"""The C-ABI library: it loads, and exports every symbol include/tfep_b200.h declares."""

import ctypes
import os
import re

import pytest

from tfep_b200 import _build, _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'tfep_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tfepb_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_build.LIBPATH):
        _build.build()
    return ctypes.CDLL(_build.LIBPATH)


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} is declared in tfep_b200.h but not exported by libtfep_b200.so'


def test_abi_version_and_status_calls(lib):
    assert lib.tfepb_abi_version() == _lib.ABI_VERSION
    lib.tfepb_last_error.restype = ctypes.c_char_p
    assert lib.tfepb_last_error() is not None
    lib.tfepb_lse_workspace_bytes.restype = ctypes.c_int64
    assert lib.tfepb_lse_workspace_bytes() > 0


def test_host_side_mt19937_seeding(lib):
    """tfepb_mt19937_seed is pure host code: compare with the oracle's init_genrand restatement."""
    import numpy as np
    from oracle.analysis_oracle import Mt19937
    st = (ctypes.c_uint32 * 625)()
    assert lib.tfepb_mt19937_seed(ctypes.c_uint32(1234), st) == 0
    assert np.array_equal(np.frombuffer(st, dtype=np.uint32)[:624], Mt19937(1234).state)
    assert st[624] == 624


def test_invalid_arguments_are_reported_without_a_gpu(lib):
    lib.tfepb_masked_linear_forward.restype = ctypes.c_int32
    assert lib.tfepb_masked_linear_forward(None, None) < 0
    lib.tfepb_last_error.restype = ctypes.c_char_p
    assert b'null' in lib.tfepb_last_error()


def test_library_holds_blackwell_tensor_core_code(lib):
    """The hot kernels are sm_100a tensor-core code, not a recompiled SIMT fallback: the SASS of the built library holds
    tcgen05 MMAs (UTCHMMA), tensor-memory loads / stores (LDTM / STTM), bulk copies of the TMA engine (UBLKCP), their
    mbarrier commits (UTCBAR) and, in the fused forward kernel, packed fp32 arithmetic (FFMA2)."""
    import shutil
    import subprocess
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([exe, '-sass', '-fun', '_ZN5tfepb5fused21maf_spline_fwd_kernelILb0ELb0EEEvNS0_6ParamsE', _build.LIBPATH],
                          capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in sass or 'SM100' in sass.upper()
    for op in ('UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'FFMA2', 'MUFU.EX2'):
        assert op in sass, f'{op} missing from the fused forward kernel'
