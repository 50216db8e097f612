"""Context measurement: the reference's algorithm as plain PyTorch ops on the SAME B200 (CUDA eager) --
the oracle restatement moved to the GPU -- for cfg2 forward at B = 65536 and the n_degrees-pass inverse at B = 8192.
Test infrastructure only (imports oracle/)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import cases
from oracle import flow_oracle as fo
dev = 'cuda:0'
flows = [m for m, _ in cases.cfg_flow('cfg2')]
for m in flows:
    m.layers = [(w.to(dev), b.to(dev)) for w, b in m.layers]
    t = m.transformer
    m.transformer = fo.Spline(x0=t.x0.to(dev), xf=t.xf.to(dev), n_bins=t.n_bins, circular=t.circular)
    m.masks = [k.to(dev) for k in m.masks]
    m.mapped, m.fixed = m.mapped.to(dev), m.fixed.to(dev)


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


try:
    old = torch.get_default_device() if hasattr(torch, 'get_default_device') else None
    x = cases.cfg_input('cfg2', 65536).to(dev)
    torch.set_default_device(dev)                      # the reference creates some tensors on the default device
    with torch.no_grad():
        fo.sequential(flows, x)
        dt, (y, ld) = timed(lambda: fo.sequential(flows, x), 5)
        print(f'reference algorithm, PyTorch CUDA eager, cfg2 forward B=65536: {dt * 1e3:.1f} ms = {65536 / dt / 1e6:.2f} M samples/s')
        yb = y[:8192].contiguous()
        dt, _ = timed(lambda: fo.sequential(flows, yb, inverse=True), 1)
        print(f'reference algorithm, PyTorch CUDA eager, cfg2 inverse B=8192: {dt * 1e3:.0f} ms = {8192 / dt / 1e3:.1f} k samples/s')
finally:
    torch.set_default_device('cpu')
