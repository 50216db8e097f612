"""cfg4 (BASELINE.json): TFEP estimator + 1000-resample bootstrap over 1e8 synthetic work values, batch-sharded
over the ranks of a torchrun launch (NCCL), plus the raw kernel bandwidth of the log-sum-exp reduction.

    python scripts/bench_cfg4.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_cfg4.py

w_i ~ N(0, 1) (analytic Delta f = -0.5, reference tests/analysis/test_bootstrap.py:178-190); every rank holds the
contiguous shard [rank n/N, (rank+1) n/N) of ONE global array generated with a seeded generator.
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from tfep_b200 import _ops
from tfep_b200.analysis import distributed as D

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lo, hi = rank * n // world, (rank + 1) * n // world
# the same global array on every rank (device generator, fixed seed), then the shard
g = torch.Generator(device=dev).manual_seed(0)
w = torch.randn(n, device=dev, generator=g)[lo:hi].clone()
torch.cuda.synchronize()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps):
    fn(); sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); sync()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), out


# raw kernel: 20 back-to-back launches (host overhead amortised), one read of the shard each
ms_k, _ = timed(lambda: _ops.lse(w, -1.0), 20)
ms_e, df = timed(lambda: D.fep_estimator_sharded(w, n_total=n), 10)
t0 = time.perf_counter()
stats = D.bootstrap_statistics_sharded(w, lo, n, n_resamples=R, generator=torch.Generator().manual_seed(1), rng='philox')
sync()
dt = time.perf_counter() - t0
if rank == 0:
    q = torch.quantile(stats.double(), torch.tensor([0.025, 0.975], dtype=torch.float64, device=dev))
    print(json.dumps({
        'config': f'cfg4: fep_estimator + {R}-resample bootstrap, n={n:.3g} work values, {world} GPU(s), batch-sharded',
        'lse_kernel_ms': ms_k, 'lse_kernel_GBps_per_gpu': 4 * (hi - lo) / ms_k / 1e6,
        'estimator_ms': ms_e, 'estimator_samples_per_s': n / ms_e * 1e3, 'df': float(df), 'analytic': -0.5,
        'bootstrap_s': dt, 'bootstrap_draws_per_s': R * n / dt, 'ci95': [float(q[0]), float(q[1])],
        'bootstrap_std': float(stats.double().std())}), flush=True)
if world > 1:
    dist.destroy_process_group()
