"""Profiling driver: a few launches of the fused tensor-core inverse chain at the headline batch (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
seq, _ = cfg_flow_modules('cfg2', 'cuda:0')
for m in seq:
    m.precision = 'bf16'
x = cases.cfg_input('cfg2', 65536).to('cuda:0')
with torch.no_grad():
    y, _ = seq(x)
    for _ in range(n):
        xi, ld = seq.inverse(y)
torch.cuda.synchronize()
print('ok', float(ld.mean()))
