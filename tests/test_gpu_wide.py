"""Wide conditioners (BASELINE.json cfg3 at D = 300 and cfg5's shape at D = 192 / 300): forward parity with the CPU
oracle at the real widths, and the BLOCKED inverse sweep (tfep_b200/_blocked.py: GEMM panels between degree blocks, the
persistent sweep inside them) against the reference's inverse loop (autoregressive.py:179-229) and the plain sweep."""

import math

import pytest
import torch

from helpers import cfg_flow_modules, rel_err
from oracle import cases
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_cfg3_forward_parity_at_full_width():
    """cfg3 at its real size D = 300 (MADE 300-670-670-1500 for SOS, 300-299-299-300 for Moebius): the fp32 path and
    the fp32-class tensor-core path against the CPU oracle, all six layers."""
    seq, flows = cfg_flow_modules('cfg3', DEV)
    x = cases.cfg_input('cfg3', 96)
    with torch.no_grad():
        y_o, ld_o = fo.sequential([m for m, _ in flows], x)
        y, ld = seq(x.to(DEV))
        assert rel_err(y, y_o) < 2e-5 and rel_err(ld, ld_o) < 5e-5
        for m in seq:
            m.precision = 'bf16x6'
        y6, ld6 = seq(x.to(DEV))
        assert rel_err(y6, y_o) < 5e-5 and rel_err(ld6, ld_o) < 1e-4


@pytest.mark.parametrize('D,block', [(192, 64), (192, 50), (300, 64)])
def test_blocked_inverse_against_the_reference_loop(D, block):
    """cfg5's layer shape (non-circular 8-bin splines, ascending and descending degrees) at D = 192 / 300: hidden widths
    960 / 1500, beyond the shared-memory sweep, so MAF.inverse runs the blocked sweep.  Against the reference's
    n_degrees-pass inverse on the CPU (fp32) and through the round trip."""
    seq, flows = cfg_flow_modules('cfg5', DEV, n_layers=2, D=D)
    B = 24
    x = cases.cfg_input('cfg5', B, D=D)
    from tfep_b200 import _blocked
    with torch.no_grad():
        y_o, ld_o = fo.sequential([m for m, _ in flows], x)
        for maf in seq:
            assert _blocked.needs_blocking(maf, maf._pack())
            maf.inverse_block_degrees = block
        y, ld = seq(x.to(DEV))
        assert rel_err(y, y_o) < 2e-5 and rel_err(ld, ld_o) < 5e-5
        xi, ldi = seq.inverse(y_o.to(DEV))
        x_ref, ld_ref = fo.sequential([m for m, _ in flows], y_o, inverse=True)       # D passes per layer on the CPU
    assert all(maf._blocked is not None for maf in seq)
    assert rel_err(xi, x_ref) < 2e-4 and rel_err(ldi, ld_ref) < 2e-4
    assert rel_err(xi, x) < 2e-4 and rel_err(ldi, -ld_o) < 2e-4


def test_blocked_inverse_equals_the_plain_sweep_and_tensor_core_panels(monkeypatch):
    """On a layer the plain persistent sweep still covers (cfg2's): forcing the blocked path must give the same x and
    log-det up to summation order; the panels on the tensor cores (bf16x6) stay within the fp32-class tolerance."""
    from tfep_b200 import _blocked
    seq, _ = cfg_flow_modules('cfg2', DEV, n_layers=2)
    y = cases.cfg_input('cfg2', 500).to(DEV) * 0.9
    with torch.no_grad():
        x_plain, ld_plain = seq.inverse(y)
        monkeypatch.setattr(_blocked, 'SWEEP_MAX_UNITS', 100)
        for maf in seq:
            maf.inverse_block_degrees = 16
        x_blk, ld_blk = seq.inverse(y)
        assert all(maf._blocked is not None for maf in seq)
        for maf in seq:
            maf.precision = 'bf16x6'
        x_tc, ld_tc = seq.inverse(y)

    def circ(a, b):
        d = (a - b).abs()
        return float(torch.minimum(d, (2 * math.pi - d).abs()).max())

    assert circ(x_blk, x_plain) < 2e-5 and rel_err(ld_blk, ld_plain) < 2e-5
    assert circ(x_tc, x_plain) < 1e-4 and rel_err(ld_tc, ld_plain) < 1e-4
