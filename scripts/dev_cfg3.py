"""cfg3 training step (bench.py secondary leg) with and without the fused transformer epilogue; device timed.
usage: python scripts/dev_cfg3.py [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tfep_b200.loss import BoltzmannKLDivLoss  # noqa: E402

dev = 'cuda:0'
B3 = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
g = torch.Generator().manual_seed(100)
x3 = torch.randn(B3, 300, generator=g).to(dev)
res = {}
modes = {'fused': (True,), 'separate': (False,)}.get(os.environ.get('DEV_CFG3_MODE', ''), (False, True))
for fuse in modes:
    seq3 = bench.build_cfg3(dev)
    for m in seq3:
        m.fuse_transformer = fuse
    opt = torch.optim.AdamW(seq3.parameters(), lr=1e-4)
    loss_fn = BoltzmannKLDivLoss()

    def step3():
        opt.zero_grad(set_to_none=True)
        y, ld = seq3(x3)
        u = 0.5 * ((y - 0.5) ** 2).sum(dim=1)
        loss = loss_fn(u, ld)
        loss.backward()
        opt.step()
        return loss

    def fwd():
        with torch.no_grad():
            return seq3(x3)

    quick = os.environ.get('DEV_CFG3_QUICK') == '1'
    for name, fn, n in ((('train_step_ms', step3, 1),) if quick else (('train_step_ms', step3, 3), ('forward_only_ms', fwd, 5))):
        for _ in range(1 if quick else 2):
            out = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        res[f'{"fused" if fuse else "separate"}_{name}'] = a.elapsed_time(b) / n
    if not quick:
        res[f'{"fused" if fuse else "separate"}_loss'] = float(step3().detach())
    res[f'{"fused" if fuse else "separate"}_peak_gb'] = torch.cuda.max_memory_allocated() / 2**30
    del seq3, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
print(json.dumps(res))
