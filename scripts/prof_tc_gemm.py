"""Profiling driver: the largest product of the cfg3 training step on the general tensor-core GEMM
(SOS output layer forward: 262144 x 1500 x 670 with bias, staircase k-block ranges)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
m, n, k = 262144, 1500, 670
x = torch.randn(m, k, device=dev)
w = torch.randn(n, k, device=dev) / k ** 0.5
b = torch.randn(n, device=dev)
# staircase: n-tile j of 256 columns sees k-blocks [0, ceil((j + 1) * 256 * k / n / 64))
ranges = torch.tensor([[0, min(11, -(-((j + 1) * 256 * k // n) // 64) + 1)] for j in range(6)], dtype=torch.int32, device=dev)
ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
for _ in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    c, img = _ops.tc_gemm(ai, bi, m, n, k, c=True, bias=b, k_block_ranges=ranges)
    e.record(); torch.cuda.synchronize()
    print('ms', a.elapsed_time(e))
