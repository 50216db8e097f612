"""Profiling driver: a hidden-layer product of the cfg3 step on the general tensor-core GEMM (262144 x 672 x 672, bias,
ELU epilogue, bf16 row image as the only output, staircase k-block ranges ~ the autoregressive mask)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
m, n, k = 262144, 672, 672
x = torch.randn(m, k, device=dev)
w = torch.randn(n, k, device=dev) / k ** 0.5
b = torch.randn(n, device=dev)
tiles_n, kblocks = -(-n // 256), -(-k // 64)
ranges = torch.tensor([[0, min(kblocks, -(-((j + 1) * 256 * k // n) // 64))] for j in range(tiles_n)], dtype=torch.int32, device=dev)
ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
for _ in range(4):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    c, img = _ops.tc_gemm(ai, bi, m, n, k, c=None, bias=b, activation=_ops.ACT_ELU, out_image=True, k_block_ranges=ranges)
    e.record(); torch.cuda.synchronize()
    print('ms', a.elapsed_time(e), 'k-block ranges', ranges.tolist())
