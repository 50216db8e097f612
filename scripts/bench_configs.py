"""Secondary measurements on the other BASELINE.json configurations (not the bench.py headline):
cfg1 (affine, D=66), cfg2 inverse, cfg3 training step (SOS + Moebius, D=300), cfg4 estimator + bootstrap.
Prints one JSON line per measurement; CUDA-event timing after warm-up."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import cfg_flow_modules
from oracle import cases

dev = 'cuda:0'


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


def out(**kw):
    print(json.dumps(kw), flush=True)


which = sys.argv[1:] or ['cfg1', 'cfg2inv', 'cfg2mix', 'cfg3', 'cfg4']
if 'cfg1' in which:
    seq, _ = cfg_flow_modules('cfg1', dev)
    x = cases.cfg_input('cfg1', 1024).to(dev)
    with torch.no_grad():
        ms = timed(lambda: seq(x), 20, 5)
        y, _ = seq(x)
        msi = timed(lambda: seq.inverse(y), 5, 2)
    out(config='cfg1 2xMAF affine D=66 B=1024', forward_ms=ms, forward_samples_per_s=1024 / ms * 1e3, inverse_ms=msi,
        inverse_samples_per_s=1024 / msi * 1e3)
if 'cfg2inv' in which:
    seq, _ = cfg_flow_modules('cfg2', dev)
    x = cases.cfg_input('cfg2', 65536).to(dev)
    with torch.no_grad():
        y, ld = seq(x)
        msi = timed(lambda: seq.inverse(y), 3, 1)
        xi, ldi = seq.inverse(y)
    d = (xi - x).abs()
    err = float(torch.minimum(d, (2 * torch.pi - d).abs()).max())
    out(config='cfg2 4xMAF spline D=66 B=65536 inverse (degree sweep, fp32)', inverse_ms=msi,
        inverse_samples_per_s=65536 / msi * 1e3, round_trip_max_err=err, logdet_cancel=float((ld + ldi).abs().max()))
if 'cfg2mix' in which:
    # BASELINE.json cfg2 "dihedral / Cartesian mix": MixedTransformer(circular spline on 22 torsions, ordinary
    # spline on 44 Cartesians), 4 layers, tensor-core chain kernels (generic spline epilogue)
    seq, _ = cfg_flow_modules('cfg2mix', dev)
    for m in seq:
        m.precision = 'bf16'
    x = cases.cfg_input('cfg2mix', 65536).to(dev)
    with torch.no_grad():
        ms = timed(lambda: seq(x), 20, 5)
        y, ld = seq(x)
        msi = timed(lambda: seq.inverse(y), 5, 2)
        xi, ldi = seq.inverse(y)
    out(config='cfg2mix 4xMAF Mixed(circular + ordinary spline) D=66 B=65536, bf16 tensor-core chain kernels',
        forward_ms=ms, forward_samples_per_s=65536 / ms * 1e3, inverse_ms=msi, inverse_samples_per_s=65536 / msi * 1e3,
        round_trip_median_err=float((xi - x).abs().max(dim=1).values.median()))
if 'cfg2mixemb' in which:
    # the shape of the reference's MixedMAFMap (app/mixedmaf.py:341-353): the mix above with the 22 torsions lifted to
    # (cos, sin) in front of the conditioner; hidden widths 334 so that the layers fit the fused kernels
    seq, _ = cfg_flow_modules('cfg2mixemb', dev, hidden_layers=[334, 334])
    for m in seq:
        m.precision = 'bf16'
    x = cases.cfg_input('cfg2mixemb', 65536).to(dev)
    with torch.no_grad():
        ms = timed(lambda: seq(x), 20, 5)
        y, ld = seq(x)
        msi = timed(lambda: seq.inverse(y), 5, 2)
    out(config='cfg2mixemb 4xMAF Mixed splines + PeriodicEmbedding (88 conditioner inputs, hidden 334) D=66 B=65536, bf16 chain kernels',
        forward_ms=ms, forward_samples_per_s=65536 / ms * 1e3, inverse_ms=msi, inverse_samples_per_s=65536 / msi * 1e3)
if 'cfg3' in which:
    B = 262144
    for precision in ('fp32', 'bf16'):
        seq, _ = cfg_flow_modules('cfg3', dev)
        for m in seq:
            m.precision = precision
        x = cases.cfg_input('cfg3', B).to(dev)
        opt = torch.optim.AdamW(seq.parameters(), lr=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            y, ld = seq(x)
            u = 0.5 * ((y - 0.5) ** 2).sum(dim=1)             # harmonic target potential
            loss = (u - ld).mean()
            loss.backward()
            opt.step()
            return loss
        ms = timed(step, 3, 1)
        kind = 'exact fp32 FFMA GEMMs' if precision == 'fp32' else 'tcgen05 GEMMs (bf16 operands, fp32 accumulation) forward and backward'
        out(config=f'cfg3 6xMAF (3 SOS + 3 Moebius) D=300 B=262144 training step (fwd+bwd+AdamW), {kind}', step_ms=ms,
            samples_per_s=B / ms * 1e3, loss=float(step().detach()))
        del seq, opt, x
        torch.cuda.empty_cache()
if 'cfg4' in which:
    from tfep_b200.analysis import bootstrap, fep_estimator
    n = 100_000_000
    g = torch.Generator(device=dev).manual_seed(0)
    w = torch.randn(n, device=dev, generator=g)
    ms = timed(lambda: fep_estimator(w), 10, 3)
    df = float(fep_estimator(w))
    out(config='cfg4 fep_estimator n=1e8', ms=ms, samples_per_s=n / ms * 1e3, gbytes_per_s=4 * n / ms / 1e6, df=df,
        analytic=-0.5)
    t0 = time.perf_counter()
    r = bootstrap(w, fep_estimator, n_resamples=1000, generator=torch.Generator().manual_seed(1), rng='philox')
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out(config='cfg4 bootstrap n=1e8 R=1000 (philox)', seconds=dt, draws_per_s=1000 * n / dt,
        ci=[float(r['confidence_interval']['low']), float(r['confidence_interval']['high'])], std=float(r['standard_deviation']))
    n2, R2 = 1_000_000, 100
    t0 = time.perf_counter()
    r = bootstrap(w[:n2], fep_estimator, n_resamples=R2, generator=torch.Generator().manual_seed(1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out(config='cfg4 bootstrap n=1e6 R=100 (exact MT19937 index stream)', seconds=dt, draws_per_s=R2 * n2 / dt,
        mean=float(r['mean']))
