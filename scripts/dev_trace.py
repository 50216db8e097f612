"""Development aid: clock64() timeline of CTA 0 of the fused kernel (debug_mode bit 16).

usage: python scripts/dev_trace.py [extra debug bits] [n_layers] [max_cycles]"""
import os, sys
os.environ['TFEPB_FUSED_DEBUG_MODE'] = str(16 | int(sys.argv[1]) if len(sys.argv) > 1 else 16)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
from tfep_b200 import _fused
dev = 'cuda:0'
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 140000
seq, _ = cfg_flow_modules('cfg2', dev, n_layers=nl)
pairs = []
for maf in seq:
    maf._fused = _fused.FusedSplinePlan(maf)
    pairs.append((maf._fused, maf))
x = cases.cfg_input('cfg2', 65536).to(dev)
ROLES = 4
with torch.no_grad():
    for _ in range(2):
        dbg = torch.zeros(ROLES * 800 * 2 + 16, dtype=torch.float32, device=dev)
        _fused.run_chain(pairs, x, debug_params=dbg)
    torch.cuda.synchronize()
t = dbg.cpu().view(torch.int64)[:ROLES * 800].reshape(ROLES, 400, 2)
ev = []
for role in range(ROLES):
    for k in range(400):
        if t[role, k, 0] != 0:
            ev.append((int(t[role, k, 0]), role, int(t[role, k, 1])))
ev.sort()
t0 = ev[0][0]
names = {0: 'PROD', 1: 'MMA ', 2: 'EPI ', 3: 'STOR'}
last = {r: t0 for r in range(ROLES)}
for ts, role, tag in ev:
    if ts - t0 > limit:
        break
    print(f'{ts - t0:8d} (+{ts - last[role]:6d}) {names[role]} {tag}')
    last[role] = ts
