import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import to_maf
from oracle import cases
dev = 'cuda:0'
for name, case in cases.maf_cases(torch.float32).items():
    _, sd = cases.build_oracle(case, torch.float32)
    maf = to_maf(case, sd, dev, torch.float32)
    x = case['x'].to(dev)
    try:
        with torch.no_grad():
            y32, ld32 = maf(x)
            maf.precision = 'bf16'
            y, ld = maf(x)
            fused = maf._fused_plan() is not None
            msg = f'fwd dy {float((y - y32).abs().max()):.2e} dld {float((ld - ld32).abs().max()):.2e}'
            if case['invertible']:
                xi, ldi = maf.inverse(y)
                msg += f' rt {float((xi - x).abs().max()):.2e}'
        print(f'{name:28s} fused={fused!s:5s} {msg}  why={getattr(maf, "_fused_why", None)}')
    except Exception as e:
        print(f'{name:28s} ERROR {type(e).__name__}: {str(e)[:150]}')
