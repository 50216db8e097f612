"""Development aid: rate of the exact MT19937 index stream (jump-ahead sub-streams) vs request size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
st = _ops.mt19937_seed(1).to(dev)
idx = torch.empty(1 << 29, dtype=torch.int32, device=dev)
for lg in (22, 24, 26, 28, 29):
    n = 1 << lg
    for streams in (None, 148, 592, 1184, 2368):
        if streams is not None and n // streams < 624:
            continue
        _ops.mt19937_indices(st, n, 100000000, out=idx, n_streams=streams)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            _ops.mt19937_indices(st, n, 100000000, out=idx, n_streams=streams)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f'2^{lg} draws, streams {streams}: {ms:.3f} ms  {n / ms / 1e6:.1f} G draws/s', flush=True)
