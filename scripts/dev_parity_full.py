"""Development aid: error distribution of every precision of the cfg2 forward at the full batch against the CPU oracle
(fp32, = the reference bit for bit) and its fp64 evaluation.  usage: python scripts/dev_parity_full.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import math
import torch
from helpers import cfg_flow_modules
from oracle import cases
from oracle import flow_oracle as fo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = 'cuda:0'
seq, flows = cfg_flow_modules('cfg2', dev)
x = cases.cfg_input('cfg2', B)
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
with torch.no_grad():
    y32, ld32 = fo.sequential([m for m, _ in flows], x)
print(f'oracle fp32: {time.perf_counter() - t0:.1f} s', flush=True)
flows64 = cases.cfg_flow('cfg2', torch.float64)
with torch.no_grad():
    y64, ld64 = fo.sequential([m for m, _ in flows64], x.double())


def circ(a, b):
    d = (a.double().cpu() - b.double().cpu()).abs()
    return torch.minimum(d, (2 * math.pi - d).abs())


def report(tag, y, ld, yr, ldr):
    dy = (circ(y, yr) / (1 + yr.abs().double())).max(dim=1).values
    dl = (ld.double().cpu() - ldr.double()).abs() / (1 + ldr.abs().double())
    q = lambda t, p: float(torch.quantile(t, p))
    print(f'{tag:24s} y: med {q(dy, .5):.2e} p99 {q(dy, .99):.2e} p99.9 {q(dy, .999):.2e} max {float(dy.max()):.2e} frac<=1e-5 {float((dy <= 1e-5).double().mean()):.5f}'
          f' | ld: med {q(dl, .5):.2e} p99 {q(dl, .99):.2e} p99.9 {q(dl, .999):.2e} max {float(dl.max()):.2e} frac<=1e-5 {float((dl <= 1e-5).double().mean()):.5f}', flush=True)


report('oracle fp32 vs fp64', y32, ld32, y64, ld64)
xd = x.to(dev)
for prec in ('fp32', 'bf16x6', 'bf16x3', 'bf16'):
    for m in seq:
        m.precision = prec
    with torch.no_grad():
        for _ in range(2):
            y, ld = seq(xd)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            y, ld = seq(xd)
        b.record()
        torch.cuda.synchronize()
    print(f'--- precision {prec}: {a.elapsed_time(b) / 3:.3f} ms per 4-layer forward of {B} samples')
    report(f'{prec} vs oracle fp32', y, ld, y32, ld32)
    report(f'{prec} vs oracle fp64', y, ld, y64, ld64)
