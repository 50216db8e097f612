"""Development check: fused bf16 kernels with a PeriodicEmbedding (odd / even numbers of plain features)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
DEV = 'cuda:0'
for cfg, D in (('cfg2mixemb', 48), ('cfg2mixemb', 47), ('cfg2mixemb', 46), ('cfg2mixemb', 33), ('cfg2mix', 12), ('cfg2mix', 5),
               ('cfg2mix', 3)):
    seq, flows = cfg_flow_modules(cfg, DEV, n_layers=2, D=D)
    x = cases.cfg_input(cfg, 1000, D=D).to(DEV)
    period = torch.full((D,), float('inf'))
    period[[f for f in range(D) if f % 3 == 2]] = 2 * math.pi
    def dist(a, b):
        d = (a.double().cpu() - b.double().cpu()).abs()
        return torch.minimum(d, (period - d).abs())
    with torch.no_grad():
        for maf in seq:
            maf.precision = 'fp32'
            y32, ld32 = maf(x)
            x32, _ = maf.inverse(y32)
            maf.precision = 'bf16'
            if maf._fused_plan() is None:
                print(D, 'not fused:', maf._fused_why)
                continue
            y, ld = maf(x)
            xi, ldi = maf.inverse(y)
            print(D, 'fwd', float(dist(y, y32).max()), float(dist(y, y32).mean()), float((ld - ld32).abs().mean()),
                  'inv', float(dist(xi, x).max(dim=1).values.median()), float((ld + ldi).abs().median()),
                  'fp32 rt', float(dist(x32, x).max()))
