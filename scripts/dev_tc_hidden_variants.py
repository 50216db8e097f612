"""Development aid: where does the time of a hidden-layer product on the general GEMM go?  Variants of scripts/prof_tc_hidden.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
m = 262144


def run(label, n, k, **kw):
    x = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev)
    ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
    if kw.pop('no_bias', False):
        b = None
    ts = []
    for it in range(7):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _ops.tc_gemm(ai, bi, m, n, k, bias=b, **kw)
        e.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(e))
    tiles = (m // 128) * (-(-n // 256)) / 148
    cyc = min(ts) * 1e-3 * 1.965e9 / tiles
    print(f'{label:58s} n={n:4d} k={k:4d}: min {min(ts):.4f} ms = {cyc / 1e3:5.2f} k cycles per tile ({-(-k // 64)} k-blocks)', flush=True)


E, N0 = _ops.ACT_ELU, _ops.ACT_NONE
for k in (64, 128, 256, 384, 512, 672):
    run('ELU, image out', 672, k, c=None, activation=E, out_image=True)
for k in (64, 672):
    run('no activation, image out', 672, k, c=None, activation=N0, out_image=True)
    run('no activation, no bias, image out', 672, k, c=None, activation=N0, out_image=True, no_bias=True)
    run('ELU, fp32 C out (no image)', 672, k, c=True, activation=E, out_image=False)
    run('no activation, fp32 C out', 672, k, c=True, activation=N0, out_image=False)
for n in (256, 512, 768):
    run('ELU, image out', n, 672, c=None, activation=E, out_image=True)
