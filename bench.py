#!/usr/bin/env python
"""Benchmark of the MAF forward + log|det J| hot path (BASELINE.json metric) on 1..N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision bf16|bf16x6|bf16x3|fp32] [--skip-secondary]

One "step" = one forward pass of the BASELINE.json configuration cfg2 (4 x MAF, MADE conditioner 66->328->328->1650,
circular neural spline K=8, D=66) over one batch of 65536 synthetic samples per GPU (weak scaling: contiguous
batch shards, no data-path collective).  Rank 0 prints ONE JSON line (contract in the task description):

  value      samples/s, inputs resident in HBM, timed per step with CUDA events (L2 flushed between steps),
             max over ranks;
  e2e        the same metric through the public module API with HOST buffers: pinned-host -> device copy of x,
             forward, device -> host copy of (y, log_det_J) inside the timed region;
  roofline   dominant kernel timed alone with CUDA events vs the measured tensor peak (MEASURED_PEAKS.json);
  cpu_baseline  the oracle (CPU restatement of the reference's PyTorch path, bit-identical to it) on the box's
             host cores, bounded sample; its outputs double as the parity check of the GPU arms (`parity`);
  parity_path   the same forward at fp32-class accuracy on the tensor cores (precision='bf16x6': operands split into
             three bf16 terms, six products per reduction step), timed beside the bf16 headline;
  cfg3 / cfg4   the other BASELINE.json configurations that exchange data between ranks: the D=300 training step with
             the data-parallel gradient all-reduce, and the estimator + 1000-resample bootstrap over 1e8 work values
             with the (max, sum exp) and per-resample all-reduces (secondary objects, outside the headline timing).
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


def product_state_dicts():
    """The weights the GPU arm evaluates (build_flow: torch.manual_seed(1234)), as CPU state dicts with the reference's
    key names: both arms and every leg of this file run the same parameters."""
    return [{k: v.detach().cpu() for k, v in m.state_dict().items()} for m in build_flow(torch.device('cpu'))]


def time_cpu_reference(steps, warmup, sample, state_dicts=None, keep_output=False):
    """The reference's CPU PyTorch path (oracle restatement), all host threads, no_grad."""
    from oracle import flow_oracle as fo
    mods = oracle_flow(state_dicts)
    x = cfg2_input(sample)
    torch.set_num_threads(os.cpu_count() or 1)
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = fo.sequential(mods, x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return sample, times, torch.get_num_threads(), (out if keep_output else None)


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path (oracle port, bit-identical to it) on the
    SAME configuration as the GPU arm -- batch 65536 per step, the same number of steps and warm-up steps, the same
    weights."""
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    sample, times, cores, _ = time_cpu_reference(steps, warmup, BATCH, state_dicts=product_state_dicts())
    total = sum(times)
    value = sample * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(times),
        'warmup': warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': 'CPU reference arm: every step is one forward of the full batch of 65536 '
                                                 'samples with the weights of the GPU arm (seed 1234)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} samples per step of the cfg2 forward, fp32, torch.no_grad, '
                                   f'{cores} threads; oracle port, bit-identical to the reference on CPU'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def committed_traffic(kernel_name):
    """(bytes per launch, source) of `kernel_name` from profiles/traffic.json -- dram__bytes_read.sum +
    dram__bytes_write.sum of a committed `ncu --set full` capture at the bench workload -- or (None, None)."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        entry = json.load(open(path)).get(kernel_name)
        if entry:
            return entry['dram_bytes_per_launch'], entry['source']
    return None, None


def _errors(y, ld, y_ref, ld_ref):
    """Relative errors |a - b| / (1 + |b|) against the CPU reference: y modulo the period of the circular splines."""
    dy = (y.double().cpu() - y_ref.double()).abs()
    dy = torch.minimum(dy, (2 * math.pi - dy).abs()) / (1 + y_ref.double().abs())
    dy = dy.max(dim=1).values
    dl = (ld.double().cpu() - ld_ref.double()).abs() / (1 + ld_ref.double().abs())
    q = lambda t, p_: float(torch.quantile(t, p_))
    return {'y': {'median': q(dy, 0.5), 'p999': q(dy, 0.999), 'max': float(dy.max()), 'frac_le_1e-5': float((dy <= 1e-5).double().mean())},
            'log_det_J': {'median': q(dl, 0.5), 'p999': q(dl, 0.999), 'max': float(dl.max()),
                          'frac_le_1e-5': float((dl <= 1e-5).double().mean())}}


def parity_legs(seq, x, y_ref, ld_ref, headline_precision, plan, pk_peaks, flush):
    """(parity of the headline arm, the fp32-class tensor-core path timed beside it): both against the CPU reference on
    the full batch of 65536 samples with identical weights."""
    with torch.no_grad():
        y, ld = seq(x)
    parity = {'against': 'CPU reference path (oracle port, fp32) on the same 65536 samples and weights',
              'precision': headline_precision, **_errors(y, ld, y_ref, ld_ref)}
    for maf in seq:
        maf.precision = 'bf16x6'
    try:
        with torch.no_grad():
            for _ in range(3):
                y6, ld6 = seq(x)
            ev = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                y6, ld6 = seq(x)
                b.record()
                ev.append((a, b))
            torch.cuda.synchronize()
        ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    finally:
        for maf in seq:
            maf.precision = headline_precision
    tflops = 2.0 * plan.masked_macs * len(seq) * BATCH / (ms * 1e-3) / 1e12
    path = {'precision': 'bf16x6', 'value': BATCH / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
            'kernels': 'per layer: tc_pack (split images of x), 3 x tc_gemm_kernel<3> (six tcgen05 products per k-step, fp32 '
                       'accumulation, ELU / split images of the activations in the epilogue), spline kernel',
            'frac': tflops / pk_peaks['bf16_tflops'],
            'frac_note': 'algorithmic masked FLOPs (one product per MAC) over the bf16 burst peak; the tensor cores issue 6x that',
            **_errors(y6, ld6, y_ref, ld_ref)}
    return parity, path


def build_cfg3(dev):
    """BASELINE.json cfg3: 6 MAF layers, D = 300 -- three SOS-polynomial layers and three Moebius layers (3-vectors,
    degrees repeated per atom), ascending / descending degrees alternating."""
    from tfep_b200.nn.conditioners.made import generate_degrees
    from tfep_b200.nn.flows import MAF, SequentialFlow
    from tfep_b200.nn.transformers import MoebiusTransformer, SOSPolynomialTransformer
    torch.manual_seed(4321)
    mafs = []
    for l in range(6):
        order = 'ascending' if l % 2 == 0 else 'descending'
        if l < 3:
            mafs.append(MAF(generate_degrees(300, order=order), SOSPolynomialTransformer(2), initialize_identity=False,
                            precision='bf16'))
        else:
            mafs.append(MAF(generate_degrees(300, order=order, repeats=3), MoebiusTransformer(dimension=3),
                            initialize_identity=False, precision='bf16'))
    return SequentialFlow(*mafs).to(dev)


def secondary_legs(dev, rank, world):
    """cfg3 and cfg4 of BASELINE.json on the same ranks: the two paths that DO exchange data.  Device
    timed, max over ranks; every rank must call this."""
    import torch.distributed as dist
    from tfep_b200.analysis import distributed as D
    from tfep_b200.loss import BoltzmannKLDivLoss
    from tfep_b200.utils.data_parallel import allreduce_gradients, broadcast_parameters

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {}
    # ---- cfg3: training step, batch 262144 per GPU, gradient all-reduce over the ranks ----
    B3 = 262144
    seq3 = build_cfg3(dev)
    broadcast_parameters(seq3)
    g = torch.Generator().manual_seed(100 + rank)
    x3 = torch.randn(B3, 300, generator=g).to(dev)
    opt = torch.optim.AdamW(seq3.parameters(), lr=1e-4)
    loss_fn = BoltzmannKLDivLoss()
    n_grad = sum(p.numel() for p in seq3.parameters())

    def step3():
        opt.zero_grad(set_to_none=True)
        y, ld = seq3(x3)
        u = 0.5 * ((y - 0.5) ** 2).sum(dim=1)                 # harmonic target potential (k = 1, mu = 0.5)
        loss = loss_fn(u, ld)
        loss.backward()
        allreduce_gradients(seq3)                             # ONE flat NCCL all-reduce (no-op at N = 1)
        opt.step()
        return loss

    for _ in range(2):
        step3()
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        loss = step3()
    b.record()
    sync()
    ms3 = max_over_ranks(a.elapsed_time(b) / 3)
    out['cfg3'] = {'workload': 'cfg3: 6xMAF (3 SOS + 3 Moebius) D=300, batch 262144 per GPU, training step = forward + '
                               'BoltzmannKLDivLoss + backward + gradient all-reduce + AdamW',
                   'precision': 'bf16 (tcgen05 GEMMs forward and backward, fp32 accumulation)', 'ms_per_step': ms3,
                   'samples_per_s': B3 * world / (ms3 * 1e-3), 'gradient_allreduce_elements': n_grad if world > 1 else 0,
                   'loss': float(loss.detach())}
    del seq3, opt, x3, loss
    torch.cuda.empty_cache()

    # ---- cfg2 as a TRAINING step (the headline flow under autograd): spline transformer + VJP in the GEMM epilogue ----
    seq2 = build_flow(dev)
    for m in seq2:
        m.precision = 'bf16'
    broadcast_parameters(seq2)
    x2 = cfg2_input(BATCH, seed=200 + rank).to(dev)
    opt2 = torch.optim.AdamW(seq2.parameters(), lr=1e-4)

    def step2():
        opt2.zero_grad(set_to_none=True)
        y, ld = seq2(x2)
        loss = loss_fn(0.5 * (y * y).sum(dim=1), ld)
        loss.backward()
        allreduce_gradients(seq2)
        opt2.step()
        return loss

    for _ in range(2):
        step2()
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        loss2 = step2()
    b.record()
    sync()
    ms2 = max_over_ranks(a.elapsed_time(b) / 5)
    out['cfg2_train'] = {'workload': 'cfg2 (4xMAF circular spline K=8, D=66), batch 65536 per GPU, training step = forward + '
                                     'BoltzmannKLDivLoss + backward + gradient all-reduce + AdamW',
                         'precision': 'bf16 (tcgen05 GEMMs, spline transformer and its VJP in the epilogue of the output product)',
                         'ms_per_step': ms2, 'samples_per_s': BATCH * world / (ms2 * 1e-3), 'loss': float(loss2.detach())}
    del seq2, opt2, x2, loss2
    torch.cuda.empty_cache()

    # ---- cfg4: estimator + 1000-resample bootstrap over 1e8 work values, contiguous shards ----
    n = 100_000_000
    lo, hi = rank * n // world, (rank + 1) * n // world
    gen = torch.Generator(device=dev).manual_seed(0)          # the same global array on every rank, then the shard
    w = torch.randn(n, device=dev, generator=gen)[lo:hi].clone()
    for _ in range(3):
        df = D.fep_estimator_sharded(w, n_total=n)
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        df = D.fep_estimator_sharded(w, n_total=n)                       # lse kernel + all-gather of (max, sum exp) pairs
    b.record()
    sync()
    ms_e = max_over_ranks(a.elapsed_time(b) / 10)
    sync()
    t0 = time.perf_counter()
    stats = D.bootstrap_statistics_sharded(w, lo, n, n_resamples=1000, generator=torch.Generator().manual_seed(1), rng='philox')
    sync()
    s_b = max_over_ranks(time.perf_counter() - t0)
    # the reference's own index stream (MT19937, bit-exact indices; jump-ahead sub-streams on all SMs): 20 resamples of
    # the full 1e8 samples -- every rank walks the whole stream and sums the draws that fall into its shard
    sync()
    t0 = time.perf_counter()
    stats_mt = D.bootstrap_statistics_sharded(w, lo, n, n_resamples=20, generator=torch.Generator().manual_seed(1))
    sync()
    s_mt = max_over_ranks(time.perf_counter() - t0)
    q = torch.quantile(stats.double(), torch.tensor([0.025, 0.975], dtype=torch.float64, device=dev))
    out['cfg4'] = {'workload': 'cfg4: fep_estimator + 1000-resample bootstrap over 1e8 work values ~ N(0,1), contiguous shards',
                   'estimator_ms': ms_e, 'estimator_samples_per_s': n / (ms_e * 1e-3),
                   'estimator_hbm_frac': 4.0 * (hi - lo) / (ms_e * 1e-3) / 1e9 / peaks()['hbm_gbs'],
                   'delta_f': float(df), 'delta_f_analytic': -0.5,
                   'bootstrap_s': s_b, 'bootstrap_draws_per_s': 1000.0 * n / s_b, 'rng': 'philox (stratified over L2-sized cells)',
                   'ci95': [float(q[0]), float(q[1])],
                   'collectives': 'all-gather of one (max, sum exp) pair per rank; MAX + SUM all-reduce of 1000 doubles',
                   'mt19937_exact_stream': {'n': n, 'n_resamples': 20, 'seconds': s_mt,
                                            'draws_per_s': 20.0 * n / s_mt, 'mean': float(stats_mt.double().mean())}}
    return out


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
    # tfep_b200.utils.host_pipeline.HostPipeline = H2D copy of x -> flow -> on-device consumer -> D2H copy of the
    # step's result.  The result a (T)FEP run consumes is the generalized work of every sample (4 bytes) and its
    # estimator partial (FEPWorkConsumer: u_target(y) - log|det J| with the synthetic harmonic target potential of
    # BASELINE.json), so that is what comes back; `e2e_full_outputs` below is the same pipeline returning y and
    # log_det_J in full (what feeds an external potential engine).
    from tfep_b200.utils.host_pipeline import FEPWorkConsumer, HostPipeline, harmonic_potential

    def time_pipeline(pipe):
        # one CUDA-graph launch per step: upload of x from pinned host memory, the chain kernel (+ consumer), download
        # of the result; three steps are in flight on three streams, so the copies of neighbouring steps overlap with
        # the kernels -- every step still moves its own inputs and outputs inside the timed region
        # Timed: from the completion of the last warm-up step to the completion of the K-th step after it (events on
        # the streams those steps run on), i.e. K steps of the running pipeline, not its fill and drain.
        pipe.step_graph(x_host)
        pipe.join()
        barrier()
        for _ in range(max(3, args.warmup)):
            pipe.step_graph(x_host)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(pipe.last_stream)
        for _ in range(args.steps):
            pipe.step_graph(x_host)
        b.record(pipe.last_stream)
        pipe.join()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    pipe = HostPipeline(seq, BATCH, 66, dev, n_chunks=1, depth=3, consumer=FEPWorkConsumer(harmonic_potential(1.0, 0.5)))
    e2e_ms = time_pipeline(pipe)
    work_host, partial_host = pipe.outputs_host
    e2e_d2h = work_host.numel() * work_host.element_size() + partial_host.numel() * partial_host.element_size()
    e2e_check = float(-(partial_host[0] + torch.log(partial_host[1]) - math.log(BATCH)))     # Delta f of the last step
    pipe_full = HostPipeline(seq, BATCH, 66, dev, n_chunks=1, depth=3)
    e2e_full_ms = time_pipeline(pipe_full)
    y_host, ld_host = pipe_full.y_host, pipe_full.ld_host

    # measured ceiling of ANY implementation that uploads x as fp32 from this host: all ranks copy their pinned batch
    # to the device at the same time, nothing else running
    # (one large copy per measurement: 8 batches = 138 MB, so that launch gaps between copies do not count)
    with gpu_numa_affinity(dev):
        big_host = torch.empty(8 * BATCH, 66, dtype=torch.float32).pin_memory()
    x_stage = torch.empty(8 * BATCH, 66, dtype=torch.float32, device=dev)
    for _ in range(2):
        x_stage.copy_(big_host, non_blocking=True)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        x_stage.copy_(big_host, non_blocking=True)
    b.record()
    torch.cuda.synchronize(dev)
    ms_stream = a.elapsed_time(b) / 32                       # per batch of 65536 samples, source streamed from host DRAM
    # ... and the pipeline's own situation: the SAME 17 MB pinned batch uploaded again and again (it can stay in the
    # host's last-level cache, which some hosts serve faster to the DMA engine than DRAM); the ceiling is the faster one
    one_host, one_dev = big_host[:BATCH], x_stage[:BATCH]
    for _ in range(4):
        one_dev.copy_(one_host, non_blocking=True)
    barrier()
    a.record()
    for _ in range(32):
        one_dev.copy_(one_host, non_blocking=True)
    b.record()
    torch.cuda.synchronize(dev)
    ms_resident = a.elapsed_time(b) / 32
    h2d_ms = torch.tensor([min(ms_stream, ms_resident)], dtype=torch.float64, device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(h2d_ms, op=dist.ReduceOp.MAX)
    h2d_ms = float(h2d_ms)
    del big_host
    del x_stage

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
        if not args.skip_secondary:
            secondary_legs(dev, rank, world)       # collective legs: every rank takes part
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
    traffic, traffic_source = committed_traffic(kernel.split(' ')[0])
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk_peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': achieved / pk_peaks['bf16_tflops'],
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload from the committed
                # `ncu --set full` capture named in traffic_source (null if none is committed for this kernel); the
                # algorithmic HBM bytes of the chain launch are 34.9 MB (x, y, log-det), y stays in L2 at kernel end
                'traffic': traffic, 'traffic_source': traffic_source,
                'kernel': kernel,
                'kernel_ms': k_ms, 'peak_source': pk_peaks['source'] + ', bf16 burst',
                'whole_step_frac': (2.0 * plan.masked_macs * 4 * BATCH * args.steps / (total_ms * 1e-3) / 1e12)
                / pk_peaks['bf16_tflops_sustained']}

    # CPU baseline = the reference's path on the host cores with the SAME weights, on the full batch; its outputs are
    # the parity reference of the GPU arms below
    sds = [{k: v.detach().cpu() for k, v in m.state_dict().items()} for m in seq]
    sample, times, cores, (y_ref, ld_ref) = time_cpu_reference(2, 1, BATCH, state_dicts=sds, keep_output=True)
    cpu_value = sample * len(times) / sum(times)
    x0_host = cfg2_input(BATCH)                     # the cpu sample = rank 0's shard at N = 1, the first shard otherwise
    parity, parity_path = parity_legs(seq, x0_host.to(dev), y_ref, ld_ref, args.precision, plan, pk_peaks, flush)

    secondary = {}
    if not args.skip_secondary:
        secondary = secondary_legs(dev, rank, world)

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
                'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': e2e_d2h,
                'api': 'tfep_b200.utils.host_pipeline.HostPipeline.step_graph with FEPWorkConsumer (CUDA graph per step, 3 '
                       'steps in flight): x up, per-sample generalized work + (max, sum exp) estimator partial down',
                'delta_f_last_step': e2e_check,
                # measured: every rank uploading its pinned fp32 batch at the same time, nothing else running
                'host_ceiling': {'h2d_ms_per_batch': h2d_ms, 'h2d_GBps_per_gpu': x_host.numel() * 4 / (h2d_ms * 1e-3) / 1e9,
                                 'samples_per_s': BATCH * world / (h2d_ms * 1e-3),
                                 'note': 'upper bound of any pipeline that uploads fp32 x from this host: concurrent '
                                         'pinned H2D copies on all ranks (PCIe / host memory system)'}},
        'e2e_full_outputs': {'value': BATCH * world * args.steps / (e2e_full_ms * 1e-3), 'unit': UNIT,
                             'h2d_bytes_per_step': x_host.numel() * 4,
                             'd2h_bytes_per_step': (y_host.numel() + ld_host.numel()) * 4,
                             'api': 'HostPipeline.step_graph without consumer: y and log_det_J downloaded in full'},
        'gpu_launches': launches_per_step * args.steps,
        'inverse': inv,
        'roofline': roofline,
        'cpu_baseline': {'value': cpu_value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} samples x {len(times)} passes of the same cfg2 forward with the same weights, '
                                   'fp32, no_grad'},
        'parity': parity,
        'parity_path': parity_path,
    }
    line.update(secondary)
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries ONE JSON line: everything any library prints there during the run (e.g. NCCL's version banner when the
    environment sets NCCL_DEBUG) is sent to stderr; `emit` writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['fp32', 'bf16', 'bf16x3', 'bf16x6'])
    ap.add_argument('--skip-secondary', action='store_true', help='only the cfg2 headline (no cfg3 / cfg4 legs)')
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
        # stdout carries ONE JSON line: NCCL's own log (e.g. the version banner of NCCL_DEBUG=VERSION) goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
