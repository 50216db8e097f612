"""Development aid: end-to-end (host -> device -> flow -> host) throughput of HostPipeline variants on cfg2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from tfep_b200.utils.host_pipeline import FEPWorkConsumer, HostPipeline, harmonic_potential

dev = torch.device('cuda', 0)
seq = bench.build_flow(dev).eval()
for m in seq:
    m.precision = 'bf16'
x_host = bench.cfg2_input(bench.BATCH).pin_memory()
B = bench.BATCH


def run(tag, pipe, steps=40):
    for _ in range(6):
        pipe.step_graph(x_host)
    pipe.join(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        pipe.step_graph(x_host)
    pipe.join(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(f'{tag:40s} {ms:.4f} ms/step  {B / ms * 1e3 / 1e6:.1f} M samples/s', flush=True)


for depth in (2, 3, 4, 6):
    run(f'full outputs, depth {depth}', HostPipeline(seq, B, 66, dev, n_chunks=1, depth=depth))
for depth in (2, 3, 4, 6):
    run(f'work consumer, depth {depth}', HostPipeline(seq, B, 66, dev, n_chunks=1, depth=depth,
                                                      consumer=FEPWorkConsumer(harmonic_potential())))
run('logdet only consumer, depth 4', HostPipeline(seq, B, 66, dev, n_chunks=1, depth=4, consumer=lambda x, y, ld: (ld,)))
# upload alone and kernel alone
xd = torch.empty(B, 66, device=dev)
for name, fn in (('H2D copy alone', lambda: xd.copy_(x_host, non_blocking=True)), ('chain kernel alone', lambda: seq(xd))):
    with torch.no_grad():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(40):
            fn()
        b.record(); torch.cuda.synchronize()
    print(f'{name:40s} {a.elapsed_time(b) / 40:.4f} ms', flush=True)
