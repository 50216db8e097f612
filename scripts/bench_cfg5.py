"""BASELINE.json cfg5: MAF 8 layers, non-circular NeuralSpline K = 8, D = 3000 (MADE 3000-14998-14998-75000 per layer),
batch 32768 sharded over 8 GPUs (4096 samples per GPU), MAF.inverse sampling.

    python scripts/bench_cfg5.py [--D 3000] [--batch 4096] [--precision fp32|bf16x6|bf16x3|bf16] [--block 64]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_cfg5.py ...

Batch-sharded, no collective on the data path: every rank inverts its own 4096 samples.  The chain alternates ascending /
descending degrees; to keep the construction time of the run short, the even layers share one set of random weights and
the odd layers another (the packed weights of all 8 layers, 46 GB in fp32, would fit; only two are built).  Per layer the
reference would run 3000 full conditioner passes (autoregressive.py:216-227), 4.2e12 multiply-accumulates per sample;
the blocked sweep (tfep_b200/_blocked.py) runs nnz(masks) = 6.98e8.  Prints one JSON line (rank 0)."""
import argparse, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--D', type=int, default=3000)
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--layers', type=int, default=8)
ap.add_argument('--precision', default='fp32')
ap.add_argument('--block', type=int, default=64)
args = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

from tfep_b200.nn.conditioners.made import generate_degrees
from tfep_b200.nn.flows import MAF
from tfep_b200.nn.transformers import NeuralSplineTransformer

D, B = args.D, args.batch


def build(order, seed):
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    lim = torch.full((D,), 5.0)
    maf = MAF(generate_degrees(D, order=order), NeuralSplineTransformer(x0=-lim, xf=lim, n_bins=8), initialize_identity=False,
              precision=args.precision)
    maf.inverse_block_degrees = args.block
    maf = maf.to(dev)
    with torch.no_grad():
        maf.inverse(torch.zeros(8, D, device=dev))          # plan + packed weights + block matrices
    torch.cuda.synchronize()
    return maf, time.perf_counter() - t0


mafs, build_s = [], []
for order, seed in (('ascending', 1), ('descending', 2)):
    m, s = build(order, seed)
    mafs.append(m)
    build_s.append(s)
plan = mafs[0]._pack()['plan']
macs = plan.masked_macs
g = torch.Generator().manual_seed(100 + rank)
y = torch.randn(B, D, generator=g).to(dev)


def chain(v):
    total = None
    for l in reversed(range(args.layers)):
        v, ld = mafs[l % 2].inverse(v)
        total = ld if total is None else total + ld
    return v, total


with torch.no_grad():
    chain(y)                                   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x, ld = chain(y)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    # round trip through the forward pass of the last-inverted layer (layer 0): inverse(forward(x)) == x
    y0, ld0 = mafs[0](x[:256])
    x0, ldi0 = mafs[0].inverse(y0)
    rt = float((x0 - x[:256]).abs().max())
    ldc = float((ld0 + ldi0).abs().max())
if rank == 0:
    flops = 2.0 * macs * args.layers * B * world
    print(json.dumps({
        'config': f'cfg5: {args.layers}xMAF non-circular spline K=8, D={D}, MADE {D}-{len(plan.packed_degrees[1])}-'
                  f'{len(plan.packed_degrees[2])}-{len(plan.packed_degrees[3])}, MAF.inverse, batch {B} per GPU x {world} GPU(s)',
        'precision_of_panels': args.precision, 'block_degrees': args.block, 'ms_per_chain_inverse': ms,
        'samples_per_s': B * world / (ms * 1e-3), 'masked_macs_per_sample_per_layer': macs,
        'algorithmic_tflops': flops / (ms * 1e-3) / 1e12,
        'reference_macs_per_sample_per_layer': float(D) * sum(len(plan.packed_degrees[l]) * len(plan.packed_degrees[l + 1]) for l in range(3)),
        'finite': bool(torch.isfinite(x).all() and torch.isfinite(ld).all()),
        'round_trip_max_abs_err_layer0': rt, 'logdet_cancellation_layer0': ldc,
        'build_seconds_per_layer': build_s, 'hbm_gb_allocated': torch.cuda.max_memory_allocated(dev) / 1e9}), flush=True)
if world > 1:
    dist.destroy_process_group()
