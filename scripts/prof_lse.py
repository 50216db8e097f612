import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200 import _ops
w = torch.randn(100_000_000, device='cuda:0')
for _ in range(3):
    o = _ops.lse(w, -1.0)
torch.cuda.synchronize()
print(o.tolist())
