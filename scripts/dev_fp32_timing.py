"""Development aid: exact fp32 path timings (cfg2 forward, cfg3 training step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev)
x = cases.cfg_input('cfg2', 65536).to(dev)
def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    print('cfg2 fp32 exact forward, 4 layers, B=65536: %.3f ms' % timed(lambda: seq(x)))
