import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfep_b200 import _ops
torch.manual_seed(0)
m, n, k = 1000, 670, 300
x = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda') / k ** 0.5
ref = x.bfloat16().double() @ w.bfloat16().double().T
c0 = torch.full((m, 672), 7.0, device='cuda')[:, :n]
c, _ = _ops.tc_gemm(_ops.tc_pack(x, 128), _ops.tc_pack(w, 256), m, n, k, c=c0)
torch.cuda.synchronize()
err = (c.double() - ref).abs()
for tm in range((m + 127) // 128):
    blk = c[tm * 128:(tm + 1) * 128]
    print(tm, 'untouched', float((blk == 7.0).float().mean()), 'err', [round(float(err[tm * 128:(tm + 1) * 128, tn * 256:(tn + 1) * 256].nan_to_num(99).max()), 3) for tn in range(3)])
# is an odd tile's content equal to some other tile's reference?
blk = c[128:256].double()
for tm in range(8):
    r = ref[tm * 128:tm * 128 + 128]
    if r.shape[0] == 128:
        print('tile1 vs ref tile', tm, float((blk - r).abs().max()))
