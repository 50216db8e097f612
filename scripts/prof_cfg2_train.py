"""Two cfg2 training steps (fused spline epilogue) for an ncu launch list: python scripts/prof_cfg2_train.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tfep_b200.loss import BoltzmannKLDivLoss  # noqa: E402

dev = 'cuda:0'
x = bench.cfg2_input(65536).to(dev)
seq = bench.build_flow(dev)
for m in seq:
    m.precision = 'bf16'
opt = torch.optim.AdamW(seq.parameters(), lr=1e-4)
loss_fn = BoltzmannKLDivLoss()
for _ in range(3):
    opt.zero_grad(set_to_none=True)
    y, ld = seq(x)
    loss = loss_fn(0.5 * (y * y).sum(dim=1), ld)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print('ok')
