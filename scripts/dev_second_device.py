"""Development check: every kernel family from ONE process on cuda:1 after cuda:0 (per-device kernel attributes, device guards)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from tfep_b200.analysis import bootstrap, fep_estimator
from tfep_b200.nn.flows import CenteredCentroidFlow, OrientedFlow
res = {}
for dev in ('cuda:0', 'cuda:1'):
    seq, _ = cfg_flow_modules('cfg2', dev, n_layers=2)
    x = ((torch.rand(3000, 66, generator=torch.Generator().manual_seed(0)) * 2 - 1) * math.pi * 0.999).to(dev)
    with torch.no_grad():
        y32, ld32 = seq(x)
        xi32, _ = seq.inverse(y32)
        for m in seq:
            m.precision = 'bf16'
        y, ld = seq(x)
        xi, _ = seq.inverse(y)
        wrapped = CenteredCentroidFlow(OrientedFlow(seq), space_dimension=3).to(dev)
        yw, ldw = wrapped(torch.randn(500, 72, generator=torch.Generator().manual_seed(1)).to(dev))
    xg = x[:200].clone().requires_grad_(True)
    yg, ldg = seq(xg)                       # bf16 under autograd: general tensor-core GEMM forward + backward
    (yg.sum() + ldg.sum()).backward()
    w = torch.randn(200000, generator=torch.Generator().manual_seed(2)).to(dev)
    df = fep_estimator(w)
    b = bootstrap(w, fep_estimator, n_resamples=50, generator=torch.Generator().manual_seed(3), rng='philox')
    res[dev] = [t.cpu() for t in (y32, ld32, xi32, y, ld, xi, yw, ldw, xg.grad, df, b['mean'])]
    torch.cuda.synchronize(dev)
ok = all(torch.equal(a, b) for a, b in zip(res['cuda:0'], res['cuda:1']))
print('cuda:1 results identical to cuda:0:', ok)
