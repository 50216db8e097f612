"""Bit-for-bit check of the oracle against the real reference; only where /root/reference exists."""

import warnings

import pytest

from oracle.ref_import import reference_available


@pytest.mark.skipif(not reference_available(), reason='reference tree not present (GPU box)')
def test_oracle_is_bit_identical_to_reference():
    from oracle import check_against_reference as chk
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        fails = chk.run_all()
    assert not fails, '\n'.join(fails)
