"""Development aid: end-to-end (host buffers) throughput of the HostPipeline for different chunk counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
from tfep_b200.utils.host_pipeline import HostPipeline
dev = torch.device('cuda:0')
seq, _ = cfg_flow_modules('cfg2', dev)
for m in seq:
    m.precision = 'bf16'
B = 65536
x_host = cases.cfg_input('cfg2', B).pin_memory()
# raw copy bandwidth
xd = torch.empty(B, 66, device=dev); yh = torch.empty(B, 66).pin_memory()
for name, fn in (('H2D', lambda: xd.copy_(x_host, non_blocking=True)), ('D2H', lambda: yh.copy_(xd, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f'{name}: {ms:.3f} ms per 17.3 MB -> {17.3e6 / ms / 1e6:.1f} GB/s')
for nch in (1, 2, 4, 8):
    pipe = HostPipeline(seq, B, 66, dev, n_chunks=nch)
    for _ in range(5):
        pipe(x_host, wait=False)
    pipe.join(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30):
        pipe(x_host, wait=False)
    pipe.join()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f'n_chunks={nch}: {ms:.3f} ms/step -> {B / ms * 1e3 / 1e6:.1f} M samples/s')

pipe = HostPipeline(seq, B, 66, dev, n_chunks=1)
for _ in range(5):
    pipe.step_graph(x_host)
pipe.join(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30):
    pipe.step_graph(x_host)
pipe.join()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 30
print(f'graph mode: {ms:.3f} ms/step -> {B / ms * 1e3 / 1e6:.1f} M samples/s')
with torch.no_grad():
    y, ld = seq(x_host.to(dev))
print('graph result equals direct call:', bool(torch.equal(pipe.y_host, y.cpu())), bool(torch.equal(pipe.ld_host, ld.cpu())))
