"""Small run of the kernels with hand-rolled synchronisation (mbarrier rings, inter-CTA tile flags, TMEM hand-overs) for
compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_fused.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
from tfep_b200 import _ops
from tfep_b200.loss import BoltzmannKLDivLoss

dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev, n_layers=2)
for m in seq:
    m.precision = 'bf16'
x = cases.cfg_input('cfg2', 300).to(dev)
with torch.no_grad():
    y, ld = seq(x)                      # fused forward chain: 2 layers, 3 tiles (one ragged), tile flags between layers
    xi, ldi = seq.inverse(y)            # fused inverse chain
    for m in seq:
        m.precision = 'bf16x6'
    y6, ld6 = seq(x)                    # split-precision tcgen05 GEMMs + spline kernel
    xe, lde = seq.inverse(y6)           # exact persistent sweep (staged weights, bulk copies)
loss = BoltzmannKLDivLoss()(ld, ld6, log_weights=ldi)
st = _ops.mt19937_seed(3).to(dev)
idx = _ops.mt19937_indices(st, 100003, 977, n_streams=7)
torch.cuda.synchronize()
print('ok', float(ld.mean()), float((xi - x).abs().median()), float(loss), int(idx.sum()))
