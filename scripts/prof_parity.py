"""Profiling driver: the fp32-class tensor-core path (precision='bf16x6') of cfg2 at the headline batch (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
seq, _ = cfg_flow_modules('cfg2', 'cuda:0')
for m in seq:
    m.precision = 'bf16x6'
x = cases.cfg_input('cfg2', 65536).to('cuda:0')
with torch.no_grad():
    for _ in range(2):
        y, ld = seq(x)
torch.cuda.synchronize()
print('ok', float(ld.mean()))
