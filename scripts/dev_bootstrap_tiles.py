"""Development timing: stratified Philox bootstrap (multinomial counts per cell) as a function of the cell size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfep_b200.analysis import bootstrap, fep_estimator
mod = sys.modules['tfep_b200.analysis.bootstrap']
dev = 'cuda:0'
n, R = 100_000_000, 400
w = torch.randn(n, generator=torch.Generator().manual_seed(0)).to(dev)
for mb in (400, 128, 96, 64, 48, 32, 24, 16, 8):
    mod.L2_TILE_ENTRIES = mb * 1024 * 1024 // 4
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = bootstrap(w, fep_estimator, n_resamples=R, generator=torch.Generator().manual_seed(1), rng='philox')
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f'cell {mb:4d} MB: {dt * 1000 / R:6.3f} s per 1000 resamples; mean {float(r["mean"]):.6f} std {float(r["standard_deviation"]):.2e}', flush=True)
