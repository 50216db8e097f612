"""Development check of the fused bf16 inverse kernel."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seq, _ = cfg_flow_modules('cfg2', dev, n_layers=nl)
x = cases.cfg_input('cfg2', B).to(dev)
def circ(a, b):
    d = (a.double() - b.double()).abs()
    return torch.minimum(d, (2 * math.pi - d).abs())
with torch.no_grad():
    for li, maf in enumerate(seq):
        maf.precision = 'bf16'
        y, ld = maf(x)
        xi, ldi = maf.inverse(y)
        torch.cuda.synchronize()
        err = int(maf._fused._tables(torch.device(dev))['err'].item())
        maf.precision = 'fp32'
        xe, lde = maf.inverse(y)
        d = circ(xi, x)
        print(f'layer {li}: watchdog={err} round trip max {d.max():.3e} median(rowmax) {d.max(dim=1).values.median():.3e} '
              f'frac<1e-3 {(d.max(dim=1).values < 1e-3).float().mean():.3f} | vs exact inverse max {circ(xi, xe).max():.3e} '
              f'mean {circ(xi, xe).mean():.3e} | ld+ldi median {(ld + ldi).abs().median():.3e} ldi vs exact mean {(ldi - lde).abs().mean():.3e}')
        print('   first row x ', x[0, :6].tolist()); print('   first row xi', xi[0, :6].tolist())
    for maf in seq:
        maf.precision = 'bf16'
    xb = cases.cfg_input('cfg2', 65536).to(dev)
    yb, _ = seq(xb)
    for _ in range(2):
        seq.inverse(yb)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); seq.inverse(yb); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(f'{len(seq)} layers inverse, B=65536: median {ms[2]:.3f} ms -> {65536 / ms[2] * 1e3 / 1e6:.2f} M samples/s')
