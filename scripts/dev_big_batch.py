"""Development check: the bf16 chain kernels on a batch of 1 000 003 samples (finite, tile-position invariant, round trip)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, math
from helpers import cfg_flow_modules
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev)
for m in seq: m.precision = 'bf16'
B = 1_000_003
x = ((torch.rand(B, 66, generator=torch.Generator().manual_seed(0)) * 2 - 1) * math.pi * 0.999).to(dev)
with torch.no_grad():
    y, ld = seq(x)
    ys, lds = seq(x[500_000:500_000 + 4099])
    xi, ldi = seq.inverse(y)
    torch.cuda.synchronize()
print('finite', bool(torch.isfinite(y).all()), bool(torch.isfinite(ld).all()), 'slice equal', torch.equal(ys, y[500_000:500_000 + 4099]), torch.equal(lds, ld[500_000:500_000+4099]))
d = (xi - x).abs(); d = torch.minimum(d, (2 * math.pi - d).abs()).max(dim=1).values
print('round trip median', float(d.median()), 'frac<1e-3', float((d < 1e-3).float().mean()), 'err flag', int(seq[0]._fused._tables(torch.device(dev))['err'].item()))
