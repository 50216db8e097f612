import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg3', dev)
for m in seq:
    m.precision = 'bf16'
x = cases.cfg_input('cfg3', 262144).to(dev)
opt = torch.optim.AdamW(seq.parameters(), lr=1e-4)
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    y, ld = seq(x)
    loss = (0.5 * ((y - 0.5) ** 2).sum(dim=1) - ld).mean()
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print('ok', float(loss))
