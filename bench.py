#!/usr/bin/env python
"""Benchmark of the MAF forward + log|det J| hot path (BASELINE.json metric) on 1..N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|bf16x3]

One "step" = one forward pass of the BASELINE.json configuration cfg2 (4 x MAF, MADE conditioner 66->328->328->1650,
circular neural spline K=8, D=66) over one batch of 65536 synthetic samples per GPU (weak scaling: contiguous
batch shards, no data-path collective).  Rank 0 prints ONE JSON line (contract in the task description):

  value      samples/s, inputs resident in HBM, timed per step with CUDA events (L2 flushed between steps),
             max over ranks;
  e2e        the same metric through the public module API with HOST buffers: pinned-host -> device copy of x,
             forward, device -> host copy of (y, log_det_J) inside the timed region;
  roofline   dominant kernel timed alone with CUDA events vs the measured tensor peak (MEASURED_PEAKS.json);
  cpu_baseline  the oracle (CPU restatement of the reference's PyTorch path, bit-identical to it) on the box's
             host cores, bounded sample.
`--impl reference` times only that CPU path (the reference is pure Python on PyTorch; /root/reference does not
exist on the GPU box, the oracle restates it bit for bit -- see oracle/check_against_reference.py).
"""

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'MAF fwd+logdet samples/s (D=66, spline)'
UNIT = 'samples/s'
WORKLOAD = 'cfg2: 4xMAF circular spline K=8, D=66, MADE 66-328-328-1650, batch 65536 per GPU'
BATCH = 65536
CPU_SAMPLE = 16384


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'], bf16_tflops_sustained=p['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for t, line in self.lines:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


N_FEATURES, N_LAYERS, N_BINS = 66, 4, 8


def cfg2_degrees(layer):
    """cfg2: layer l uses ascending degrees if l is even, descending otherwise."""
    d = torch.arange(N_FEATURES)
    return d if layer % 2 == 0 else d.flip(0)


def cfg2_input(batch, seed=0):
    """x = (rand * 2 - 1) * pi * 0.999, generated on the host from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, N_FEATURES, generator=g) * 2 - 1) * math.pi * 0.999


def build_flow(device):
    """cfg2 with this package's modules: 4 x MAF(circular spline K = 8) over 66 features, random-init weights
    (torch.manual_seed(1234), initialize_identity=False as in SURVEY.md 8d: identity init would make the transformer
    parameters input-independent)."""
    from tfep_b200.nn.flows import MAF, SequentialFlow
    from tfep_b200.nn.transformers import NeuralSplineTransformer
    torch.manual_seed(1234)
    lim = torch.full((N_FEATURES,), math.pi)
    mafs = [MAF(degrees_in=cfg2_degrees(l),
                transformer=NeuralSplineTransformer(x0=-lim, xf=lim, n_bins=N_BINS, circular=True),
                hidden_layers=2, weight_norm=True, initialize_identity=False) for l in range(N_LAYERS)]
    return SequentialFlow(*mafs).to(device)


def oracle_flow(state_dicts=None):
    """The same cfg2 flow as a list of oracle layers (CPU restatement of the reference).  Test infrastructure: only the
    cpu_baseline leg and the reference arm come here.  ``state_dicts``: per layer, the product modules' parameters
    (reference key names), so that both arms evaluate the same weights; None: the oracle's own seeded weights."""
    from oracle import cases
    from oracle import flow_oracle as fo
    if state_dicts is None:
        return [m for m, _ in cases.cfg_flow('cfg2')]
    lim = torch.full((N_FEATURES,), math.pi)
    return [fo.MafOracle(cfg2_degrees(l), fo.Spline(x0=-lim, xf=lim, n_bins=N_BINS, circular=True)).load(sd)
            for l, sd in enumerate(state_dicts)]


def time_cpu_reference(steps, warmup, sample=CPU_SAMPLE, state_dicts=None):
    """The reference's CPU PyTorch path (oracle restatement), all host threads, no_grad."""
    from oracle import flow_oracle as fo
    mods = oracle_flow(state_dicts)
    x = cfg2_input(sample)
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fo.sequential(mods, x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return sample, times, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    sample, times, cores = time_cpu_reference(steps, warmup)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(times),
        'warmup': warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': f'CPU reference arm: each step is a bounded sample of {sample} samples'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} samples per step of the cfg2 forward, fp32, torch.no_grad, '
                                   f'{cores} threads; oracle port, bit-identical to the reference on CPU'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from tfep_b200 import _ops

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    from tfep_b200.utils.host_pipeline import gpu_numa_affinity
    seq = build_flow(dev)
    seq.eval()
    for maf in seq:
        maf.precision = args.precision
    # contiguous batch shards of one global synthetic data set (seeded on the host)
    x_host = cfg2_input(BATCH * world)[rank * BATCH:(rank + 1) * BATCH].contiguous()
    with gpu_numa_affinity(dev):              # pinned input buffer on the NUMA node of this rank's GPU
        x_host = x_host.pin_memory()
    x = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step():
        with torch.no_grad():
            return seq(x)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank).__enter__()
    time.sleep(0.6)                             # let nvidia-smi start sampling
    for _ in range(args.warmup):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_load0 = time.perf_counter()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()                           # L2 flush between timed iterations (untimed)
        a.record()
        y, ld = step()
        b.record()
    torch.cuda.synchronize(dev)
    t_wall = time.perf_counter() - t_wall0
    # keep the same load running (untimed) until the 100 ms clock sampler has seen at least ~1 s of it
    while time.perf_counter() - t_load0 < 1.0:
        step()
        torch.cuda.synchronize(dev)
    t_load1 = time.perf_counter()
    clocks.__exit__()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)

    # end to end through the public API with host buffers (pinned), copies inside the timed region:
    # tfep_b200.utils.host_pipeline.HostPipeline = chunked H2D copy -> flow -> D2H copy of (y, log_det_J) on three streams
    from tfep_b200.utils.host_pipeline import HostPipeline
    pipe = HostPipeline(seq, BATCH, 66, dev, n_chunks=1, depth=3)

    def e2e_step():
        # one CUDA-graph launch per step: upload of x from pinned host memory, the chain kernel, download of
        # (y, log_det_J); three steps are in flight on three streams, so the copies of neighbouring steps overlap
        # with the kernels -- every step still moves its own inputs and outputs inside the timed region
        pipe.step_graph(x_host)

    for _ in range(max(1, args.warmup)):
        e2e_step()
    pipe.join()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        e2e_step()
    pipe.join()
    b.record()
    torch.cuda.synchronize(dev)
    y_host, ld_host = pipe.y_host, pipe.ld_host
    e2e_ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms)

    # secondary: MAF.inverse of the same configuration (cfg2 is "forward + inverse + log-det"), device resident
    inv = None
    if args.precision == 'bf16':
        with torch.no_grad():
            y_dev, _ = seq(x)
            for _ in range(2):
                seq.inverse(y_dev)
            torch.cuda.synchronize(dev)
            iev = []
            for _ in range(5):
                flush.zero_()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                x_back, ld_back = seq.inverse(y_dev)
                e_.record()
                iev.append((s_, e_))
            torch.cuda.synchronize(dev)
            inv_ms = statistics.median(s_.elapsed_time(e_) for s_, e_ in iev)
            d = (x_back - x).abs()
            d = torch.minimum(d, (2 * 3.141592653589793 - d).abs()).max(dim=1).values
            inv = {'samples_per_s': BATCH / (inv_ms * 1e-3), 'ms': inv_ms, 'kernel': 'maf_spline_inv_kernel (one launch per step)',
                   'round_trip_median_abs_err': float(d.median()), 'round_trip_frac_below_1e-3': float((d < 1e-3).float().mean())}

    if rank != 0:
        return

    pk_peaks = peaks()
    pk = seq[0]._pack()
    plan = pk['plan']
    if args.precision == 'bf16':
        # dominant (only) kernel: the fused MAF-chain kernel (all layers in ONE persistent launch), timed alone
        with torch.no_grad():
            for _ in range(3):
                seq(x)
            kev = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                seq(x)
                e.record()
                kev.append((s, e))
            torch.cuda.synchronize(dev)
        k_ms = statistics.mean(s.elapsed_time(e) for s, e in kev)
        # algorithmic: 2 x non-zeros of the three masks x layers x samples
        flops_per_launch = 2.0 * plan.masked_macs * len(seq) * BATCH
        kernel = (f'maf_spline_fwd_kernel ({len(seq)} MAF layers in one launch: 3 masked GEMMs on tcgen05 + ELU + spline '
                  'epilogue per layer)')
        launches_per_step = 1
    else:
        # dominant kernel of the exact path: the output-layer GEMM (328 -> 1650) of one MAF layer, timed alone
        with torch.no_grad():
            pw, pb = seq[0]._conditioner.packed_weights(plan)
            k_ranges, _, _ = plan.tables(dev)
            h = torch.randn(BATCH, pw[2].shape[1], device=dev)
            out = torch.empty(BATCH, pw[2].shape[0], device=dev)
            for _ in range(3):
                _ops.linear_forward(h, pw[2], pb[2], 0, k_ranges[2], out=out)
            kev = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _ops.linear_forward(h, pw[2], pb[2], 0, k_ranges[2], out=out)
                e.record()
                kev.append((s, e))
            torch.cuda.synchronize(dev)
        k_ms = statistics.mean(s.elapsed_time(e) for s, e in kev)
        flops_per_launch = 2.0 * plan.nnz[2] * BATCH           # algorithmic: 2 x non-zeros of the mask x samples
        kernel = 'gemm_kernel<float,true,true> (output layer 328->1650, fp32 FFMA, staircase-skipped)'
        launches_per_step = 16
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk_peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': achieved / pk_peaks['bf16_tflops'],
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture of the
                # bf16 chain launch at this workload (profiles/r01_ncu_fused_fwd_v8_chain_details.md): 21.08 MB read
                # (x 17.3 MB + packed weights) + 0.15 MB written -- the 17.6 MB of y / logdet were still in L2 when
                # the kernel ended; the algorithmic HBM bytes are 34.9 MB per launch
                'traffic': 21.23e6 if (args.precision == 'bf16' and BATCH == 65536) else None,
                'kernel': kernel,
                'kernel_ms': k_ms, 'peak_source': pk_peaks['source'] + ', bf16 burst',
                'whole_step_frac': (2.0 * plan.masked_macs * 4 * BATCH * args.steps / (total_ms * 1e-3) / 1e12)
                / pk_peaks['bf16_tflops_sustained']}

    sample, times, cores = time_cpu_reference(2, 1, state_dicts=[{k: v.detach().cpu() for k, v in m.state_dict().items()}
                                                                  for m in seq])
    cpu_value = sample * len(times) / sum(times)

    value = BATCH * world * args.steps / (total_ms * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'precision': args.precision, 'l2': 'flushed (256 MB memset) between timed steps',
                   'parallelism': f'batch-sharded x{world}, weights replicated, no data-path collective',
                   'wall_s_timed_region': t_wall},
        'clocks': clocks.summary(t_load0, t_load1),
        'e2e': {'value': BATCH * world * args.steps / (e2e_ms * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': (y_host.numel() + ld_host.numel()) * 4,
                'api': 'tfep_b200.utils.host_pipeline.HostPipeline.step_graph (CUDA graph per step, 3 steps in flight)'},
        'gpu_launches': launches_per_step * args.steps,
        'inverse': inv,
        'roofline': roofline,
        'cpu_baseline': {'value': cpu_value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} samples x {len(times)} passes of the same cfg2 forward, fp32, no_grad'},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['fp32', 'bf16'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
