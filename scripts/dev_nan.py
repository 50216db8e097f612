import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev='cuda:0'
B=int(sys.argv[1]) if len(sys.argv)>1 else 262144
seq, flows = cfg_flow_modules('cfg3', dev)
x = cases.cfg_input('cfg3', B).to(dev)
cur = x.clone().requires_grad_(True)
h = cur
tot = 0
for li, maf in enumerate(seq):
    y, ld = maf(h)
    print(li, type(maf._transformer).__name__, 'y nan', int(torch.isnan(y).sum()), 'ld nan', int(torch.isnan(ld).sum()), 'ld inf', int(torch.isinf(ld).sum()), 'y absmax', float(y.abs().max()), 'ld range', float(ld.min()), float(ld.max()))
    h = y
    tot = tot + ld
u = 0.5 * ((h - 0.5) ** 2).sum(dim=1)
loss = (u - tot).mean()
loss.backward()
print('loss', float(loss))
for n, p in seq.named_parameters():
    if p.grad is not None and not torch.isfinite(p.grad).all():
        print('non-finite grad', n, int((~torch.isfinite(p.grad)).sum()))
print('x grad finite', bool(torch.isfinite(cur.grad).all()))
opt = torch.optim.AdamW(seq.parameters(), lr=1e-4)
for it in range(6):
    opt.zero_grad(set_to_none=True)
    y, ld = seq(x)
    u = 0.5 * ((y - 0.5) ** 2).sum(dim=1)
    loss = (u - ld).mean()
    loss.backward()
    bad = [n for n, p in seq.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print('iter', it, 'loss', float(loss.detach()), 'nan y', int(torch.isnan(y).sum()), 'nan ld', int(torch.isnan(ld).sum()), 'inf ld', int(torch.isinf(ld).sum()), 'bad grads', bad[:3])
    opt.step()

