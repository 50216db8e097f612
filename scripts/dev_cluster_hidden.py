"""Development aid: the hidden-layer product of scripts/prof_tc_hidden.py with and without cluster mode (B shared by two CTAs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (m, n, k) in [(262144, 672, 672), (65536, 336, 336), (262144, 672, 304)]:
    x = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev)
    tiles_n, kblocks = -(-n // 256), -(-k // 64)
    ranges = torch.tensor([[0, min(kblocks, -(-((j + 1) * 256 * k // n) // 64))] for j in range(tiles_n)], dtype=torch.int32, device=dev)
    ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
    res = {}
    for cluster in (False, True, False, True):
        for rr in (ranges, None):
            ts = []
            for it in range(6):
                flush.zero_()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                c, img = _ops.tc_gemm(ai, bi, m, n, k, c=None, bias=b, activation=_ops.ACT_ELU, out_image=True, k_block_ranges=rr, cluster=cluster)
                e.record(); torch.cuda.synchronize()
                if it >= 2:
                    ts.append(a.elapsed_time(e))
            key = (cluster, rr is not None)
            if key in res:
                assert torch.equal(res[key][1], img)
            res[key] = (min(ts), img)
            print(f'{m}x{n}x{k} cluster={cluster} staircase={rr is not None}: min {min(ts):.4f} ms  median {sorted(ts)[len(ts) // 2]:.4f} ms', flush=True)
    print('images equal with / without clusters:', torch.equal(res[(False, False)][1], res[(True, False)][1]))
