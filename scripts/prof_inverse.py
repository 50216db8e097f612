"""Profiling driver: a few launches of the persistent inverse sweep (one cfg2 MAF layer) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
seq, _ = cfg_flow_modules('cfg2', 'cuda:0', n_layers=1)
y = cases.cfg_input('cfg2', B).to('cuda:0')
with torch.no_grad():
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        x, ld = seq[0].inverse(y)
        b.record()
        torch.cuda.synchronize()
        print('ms', a.elapsed_time(b))
print('ok', float(ld.mean()))
