import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases
dev='cuda:0'
seq, _ = cfg_flow_modules('cfg2', dev)
for m in seq: m.precision='bf16'
x = cases.cfg_input('cfg2', 65536).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(label, fn, n=20, do_flush=False):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ev=[]
        t0=time.perf_counter()
        for _ in range(n):
            if do_flush: flush.zero_()
            a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); ev.append((a,b))
        t_enq=time.perf_counter()-t0
        torch.cuda.synchronize()
        t1=time.perf_counter()-t0
    ms=[a.elapsed_time(b) for a,b in ev]
    print(f'{label:40s} gpu mean {sum(ms)/n:.3f} ms  min {min(ms):.3f}  cpu enqueue {1e3*t_enq/n:.3f} ms/iter wall {1e3*t1/n:.3f}')
run('seq(x)', lambda: seq(x))
run('seq(x) + flush', lambda: seq(x), do_flush=True)
run('4 fused calls', lambda: [m._fused.forward(m, x) for m in seq])
run('4 fused calls + flush', lambda: [m._fused.forward(m, x) for m in seq], do_flush=True)
run('1 fused call', lambda: seq[0]._fused.forward(seq[0], x))
run('1 fused call + flush', lambda: seq[0]._fused.forward(seq[0], x), do_flush=True)
