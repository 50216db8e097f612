import os, sys
os.environ['TFEPB_FUSED_DEBUG_MODE'] = str(16 | int(sys.argv[1]) if len(sys.argv) > 1 else 16)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
from tfep_b200 import _fused
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev, n_layers=1)
maf = seq[0]
plan = _fused.FusedSplinePlan(maf)
x = cases.cfg_input('cfg2', 65536).to(dev)
with torch.no_grad():
    for _ in range(2):
        dbg = torch.zeros(3 * 800 * 2 + 16, dtype=torch.float32, device=dev)
        plan.forward(maf, x, debug_params=dbg)
    torch.cuda.synchronize()
t = dbg.cpu().view(torch.int64)[:3 * 800 // 1].reshape(3, 400, 2)
ev = []
for role in range(3):
    for k in range(400):
        if t[role, k, 0] != 0:
            ev.append((int(t[role, k, 0]), role, int(t[role, k, 1])))
ev.sort()
t0 = ev[0][0]
names = {0: 'PROD', 1: 'MMA ', 2: 'EPI '}
last = {0: t0, 1: t0, 2: t0}
n = 0
for ts, role, tag in ev:
    if ts - t0 > 140000:
        break
    print(f'{ts - t0:8d} (+{ts - last[role]:6d}) {names[role]} {tag}')
    last[role] = ts
