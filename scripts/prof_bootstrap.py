"""Profiling driver: the bootstrap gather kernel on one 32 MB cell of the exp table (stratified Philox resampling),
100 resamples of ~8.3e6 draws each -- the launch cfg4 repeats 12 times per 1000 resamples of 1e8 draws."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
dev = 'cuda:0'
cell, R = 8 << 20, 1000
e = torch.rand(cell, device=dev)
sizes = torch.full((R,), cell, dtype=torch.int64, device=dev)
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s = _ops.bootstrap_sums(e, cell, R, cell, None, 1234, 0, sample_sizes=sizes)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print('ms', ms, 'G draws/s', R * cell / ms / 1e6)
