import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
seq, _ = cfg_flow_modules('cfg2', 'cuda:0', n_layers=1)
x = cases.cfg_input('cfg2', 65536).to('cuda:0')
with torch.no_grad():
    for _ in range(3):
        y, ld = seq(x)
torch.cuda.synchronize()
print('ok')
