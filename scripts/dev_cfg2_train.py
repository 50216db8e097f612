"""cfg2 (4 x MAF circular spline, D = 66, batch 65536) TRAINING step at precision='bf16' with and without the fused spline
epilogue of the general tensor-core GEMM, and the wide spline flow of the reference's MixedMAFMap default (hidden 381) in
inference; device timed.  python scripts/dev_cfg2_train.py"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tfep_b200.loss import BoltzmannKLDivLoss  # noqa: E402
from tfep_b200.nn.flows import MAF, SequentialFlow  # noqa: E402
from tfep_b200.nn.transformers import NeuralSplineTransformer  # noqa: E402

dev = 'cuda:0'
B = 65536
x = bench.cfg2_input(B).to(dev)
res = {}


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for fuse in (False, True):
    seq = bench.build_flow(dev)
    for m in seq:
        m.precision = 'bf16'
        m.fuse_transformer = fuse
    opt = torch.optim.AdamW(seq.parameters(), lr=1e-4)
    loss_fn = BoltzmannKLDivLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        y, ld = seq(x)
        loss = loss_fn(0.5 * (y * y).sum(dim=1), ld)
        loss.backward()
        opt.step()
        return loss

    tag = 'fused' if fuse else 'separate'
    res[f'cfg2_train_{tag}_ms'] = timed(step)
    res[f'cfg2_train_{tag}_loss'] = float(step().detach())

# wide spline flow (hidden 381 > the one-launch kernel's tensor-memory plan), inference
torch.manual_seed(5)
lim = torch.full((66,), math.pi)
wide = SequentialFlow(*[MAF(bench.cfg2_degrees(l), NeuralSplineTransformer(x0=-lim, xf=lim, n_bins=8, circular=True),
                            hidden_layers=[381, 381], initialize_identity=False, precision='bf16') for l in range(4)]).to(dev)
for fuse in (False, True):
    for m in wide:
        m.fuse_transformer = fuse
    with torch.no_grad():
        res[f'wide381_forward_{"fused" if fuse else "separate"}_ms'] = timed(lambda: wide(x))
print(json.dumps(res))
