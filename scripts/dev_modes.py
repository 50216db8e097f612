"""Development aid: time the fused chain under the kernel's debug modes (results are garbage, timing is valid).
bit 1: skip weight copies; 2: skip spline math; 4: no ex2 in the ELU of the hidden layers; 8: one k-step per block;
32: debug build, no effect."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev = 'cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev)
for m in seq:
    m.precision = 'bf16'
x = cases.cfg_input('cfg2', 65536).to(dev)
for mode in [int(a) for a in sys.argv[1:]] or [0, 32, 33, 34, 36, 38, 40, 46, 47]:
    os.environ['TFEPB_FUSED_DEBUG_MODE'] = str(mode)
    with torch.no_grad():
        for _ in range(3):
            seq(x)
        torch.cuda.synchronize()
        ev = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); seq(x); b.record(); ev.append((a, b))
        torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(f'mode {mode:3d}: 4-layer chain median {ms[5]:.3f} ms min {ms[0]:.3f} ms', flush=True)
