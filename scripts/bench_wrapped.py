"""Timing: cfg2's 4-layer bf16 chain wrapped as CenteredCentroidFlow(OrientedFlow(chain)) on 24 points (72 features ->
66 through the flow), batch 65536 -- fused pre / post kernels vs the tensor-algebra path of the wrappers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from tfep_b200.nn.flows import CenteredCentroidFlow, OrientedFlow
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev)
for m in seq:
    m.precision = 'bf16'
flow = CenteredCentroidFlow(OrientedFlow(seq, axis_point_idx=0, plane_point_idx=1), space_dimension=3, fixed_point_idx=0).to(dev)
x = torch.randn(65536, 72, generator=torch.Generator().manual_seed(0)).to(dev)


def timed(fn, n=20):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); ev.append((a, b))
        torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]


bare = timed(lambda: seq(x[:, :66].contiguous()))
fused = timed(lambda: flow(x))
os.environ['TFEPB_NO_FRAME_KERNELS'] = '1'
algebra = timed(lambda: flow(x))
print(f'chain alone {bare:.3f} ms; wrapped, fused pre/post kernels {fused:.3f} ms; wrapped, tensor algebra {algebra:.3f} ms')
