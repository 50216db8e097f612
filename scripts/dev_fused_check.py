"""Development check of the fused bf16 kernel: GEMM chain vs a bf16-emulating reference, then y / logdet."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases, flow_oracle as fo
from tfep_b200 import _fused

torch.manual_seed(0)
dev = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seq, flows = cfg_flow_modules('cfg2', dev, n_layers=nl)
x = cases.cfg_input('cfg2', B)
xd = x.to(dev)


def bf(t):
    return t.to(torch.bfloat16).double()


def emulate(maf_oracle, x):
    L2E = 1.4426950408889634
    (w1, b1), (w2, b2), (w3, b3) = [(w.double(), b.double()) for w, b in maf_oracle.layers]
    act = lambda t: torch.where(t > 0, t, L2E * (torch.exp2(t) - 1))
    a1 = act(bf(x) @ bf((w1 * L2E).float()).T + b1 * L2E)
    a2 = act(bf(a1.float()) @ bf(w2.float()).T + b2 * L2E)
    n_feat = w3.shape[0] // 25
    scale = torch.ones(w3.shape[0], dtype=torch.double)
    scale[:24 * n_feat] = L2E
    return (bf(a2.float()) @ bf((w3 * (scale / L2E)[:, None]).float()).T) / scale + b3


with torch.no_grad():
    cur, cur_o = xd, x
    for li, maf in enumerate(seq):
        plan = _fused.FusedSplinePlan(maf)
        maf._fused = plan
        dbg = torch.zeros(B, plan.n_chunks * _fused.CHUNK_N, device=dev)
        y, ld = plan.forward(maf, cur, debug_params=dbg)
        torch.cuda.synchronize()
        err = int(plan._tables(torch.device(dev))['err'].item())
        # expected params (reference order) -> packed order
        par_ref = emulate(flows[li][0], cur_o)
        rows = plan.w3_rows
        exp = torch.zeros(B, len(rows), dtype=torch.double)
        valid = rows >= 0
        exp[:, valid] = par_ref[:, rows[valid]] * plan.w3_scale[valid].double()
        got = dbg.cpu().double()
        d = (got[:, valid] - exp[:, valid]).abs()
        print(f'layer {li}: watchdog={err} params max abs err {d.max():.3e} mean {d.mean():.3e} (scale {exp.abs().mean():.3f})')
        worst = d.max(dim=0).values
        print('   worst columns:', worst.topk(5).indices.tolist(), worst.topk(5).values.tolist())
        # transformer on emulated params vs kernel outputs
        old = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
        spec = cases.as_double(flows[li][0].transformer)
        y_e, ld_e = spec.forward(cur_o.double(), par_ref)
        y_64, ld_64 = None, None
        torch.set_default_dtype(old)
        y_o, ld_o = flows[li][0].forward(cur_o)
        print(f'   y vs bf16-emulation: {(y.cpu().double()-y_e).abs().max():.3e}  ld: {(ld.cpu().double()-ld_e).abs().max():.3e}'
              f' | y vs fp32 oracle: max {(y.cpu()-y_o).abs().max():.3e} mean {(y.cpu()-y_o).abs().mean():.3e}'
              f'  ld: max {(ld.cpu()-ld_o).abs().max():.3e} mean {(ld.cpu()-ld_o).abs().mean():.3e}')
        cur, cur_o = y, y.cpu()
# timing at the headline batch
xb = cases.cfg_input('cfg2', 65536).to(dev)
with torch.no_grad():
    for maf in seq:
        maf.precision = 'bf16'
    for _ in range(3):
        seq(xb)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record(); seq(xb); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(f'{len(seq)} layers, B=65536: median {ms[len(ms)//2]:.3f} ms, min {ms[0]:.3f} ms -> {65536/ms[len(ms)//2]*1e3/1e6*4/len(seq)/4:.1f} M samples/s per {len(seq)}-layer pass')
print('done')
