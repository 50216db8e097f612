"""Generate tests/golden/* by running the REAL reference (build container only).

TEST INFRASTRUCTURE ONLY.  ``python -m oracle.make_golden`` imports
andrrizzi/tfep from /root/reference (oracle/ref_import.py), evaluates it on the
seeded cases of oracle/cases.py and stores inputs + reference outputs.  The
files are committed; the GPU box never sees the reference.  Conditioner weights
are not stored: they are regenerated from their seed (cases.seeded_state) and
guarded by a checksum stored next to the outputs.
"""

import json
import os
import sys

import numpy as np
import torch

from . import cases
from . import flow_oracle as fo
from .check_against_reference import to_reference_maf, to_reference_transformer, to_reference_wrapper
from .ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def golden_transformers(ref, dtype, tag):
    out = {}
    for name, (spec, n, x, par) in cases.transformer_cases(dtype).items():
        t = to_reference_transformer(ref, spec)
        y, ld = t(x, par)
        out[f'{name}/x'], out[f'{name}/par'], out[f'{name}/y'], out[f'{name}/ld'] = _np(x), _np(par), _np(y), _np(ld)
        if not isinstance(spec, fo.SOS):
            xi, ldi = t.inverse(y, par)
            out[f'{name}/xinv'], out[f'{name}/ldinv'] = _np(xi), _np(ldi)
        else:
            # gradient of sum(y * c) through the reference's hand-written backward (sos.py:237-268)
            xg, pg = x.clone().requires_grad_(True), par.clone().requires_grad_(True)
            c = cases.normal(tuple(x.shape), 77, dtype)
            yy, _ = t(xg, pg)
            (yy * c).sum().backward()
            out[f'{name}/gy'], out[f'{name}/gx'], out[f'{name}/gpar'] = _np(c), _np(xg.grad), _np(pg.grad)
        if isinstance(spec, fo.Spline):
            out[f'{name}/bins'] = _np(fo.spline_bins(spec, x, par))       # oracle == reference bitwise (checked)
    np.savez_compressed(os.path.join(OUT, f'transformers_{tag}.npz'), **out)


def golden_mafs(ref, dtype, tag):
    out = {}
    for name, case in cases.maf_cases(dtype).items():
        oracle, sd = cases.build_oracle(case, dtype)
        maf = to_reference_maf(ref, case, sd)
        with torch.no_grad():
            y, ld = maf(case['x'])
            out[f'{name}/x'], out[f'{name}/y'], out[f'{name}/ld'] = _np(case['x']), _np(y), _np(ld)
            out[f'{name}/checksum'] = np.float64(cases.checksum(sd))
            if case['invertible']:
                xi, ldi = maf.inverse(y)
                out[f'{name}/xinv'], out[f'{name}/ldinv'] = _np(xi), _np(ldi)
        # gradients of a scalar loss w.r.t. x and all conditioner parameters (reference autograd)
        xg = case['x'].clone().requires_grad_(True)
        y, ld = maf(xg)
        cy, cl = cases.normal(tuple(y.shape), 78, dtype), cases.normal(tuple(ld.shape), 79, dtype)
        loss = (y * cy).sum() + ((ld * cl).sum() if ld.requires_grad else 0.0)
        loss.backward()
        out[f'{name}/gx'] = _np(xg.grad)
        for k, p in maf.named_parameters():
            out[f'{name}/grad/{k}'] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, f'maf_{tag}.npz'), **out)


def golden_wrappers(ref, dtype, tag):
    """PartialFlow / CenteredCentroidFlow / OrientedFlow around a reference MAF: outputs, inverse, gradient w.r.t. x."""
    out = {}
    for name, case in cases.wrapper_cases(dtype).items():
        _, sd = cases.build_wrapper_oracle(case, dtype)
        flow = to_reference_wrapper(ref, case, sd, dtype)
        with torch.no_grad():
            y, ld = flow(case['x'].clone())
            out[f'{name}/x'], out[f'{name}/y'], out[f'{name}/ld'] = _np(case['x']), _np(y), _np(ld)
            out[f'{name}/checksum'] = np.float64(cases.checksum(sd))
            if case['invertible']:
                xi, ldi = flow.inverse(y.clone())
                out[f'{name}/xinv'], out[f'{name}/ldinv'] = _np(xi), _np(ldi)
        xg = case['x'].clone().requires_grad_(True)
        y, ld = flow(xg)
        cy, cl = cases.normal(tuple(y.shape), 78, dtype), cases.normal(tuple(ld.shape), 79, dtype)
        ((y * cy).sum() + (ld * cl).sum()).backward()
        out[f'{name}/gx'] = _np(xg.grad)
    np.savez_compressed(os.path.join(OUT, f'wrappers_{tag}.npz'), **out)


def golden_embeddings(ref):
    """FlipInvariantEmbedding / MixedEmbedding of the reference with seeded parameters: state dict, input, output,
    output degrees and the gradient of sum(out * c) w.r.t. x."""
    out = {}
    for name, (build, n, deg) in cases.embedding_cases().items():
        torch.manual_seed(40)
        emb = build(ref)
        x = cases.normal((9, n), 41)
        xg = x.clone().requires_grad_(True)
        y = emb(xg)
        c = cases.normal(tuple(y.shape), 42)
        (y * c).sum().backward()
        out[f'{name}/x'], out[f'{name}/y'], out[f'{name}/gx'] = _np(x), _np(y), _np(xg.grad)
        out[f'{name}/deg'] = _np(emb.get_degrees_out(deg))
        for k, v in emb.state_dict().items():
            out[f'{name}/sd/{k}'] = _np(v)
    np.savez_compressed(os.path.join(OUT, 'embeddings.npz'), **out)


def golden_logger(ref):
    """A log directory written by the reference's TFEPLogger (tests/golden/tfep_logs) and what it reads back."""
    import shutil
    d = os.path.join(OUT, 'tfep_logs')
    shutil.rmtree(d, ignore_errors=True)
    _, reads = cases.logger_script(ref.TFEPLogger, d)
    np.savez_compressed(os.path.join(OUT, 'tfep_logs_reads.npz'),
                        **{f'{k}/{n}': v for k, r in reads.items() for n, v in r.items()})


def golden_cfg(ref):
    """Slices of the BASELINE.json configurations, fp32 reference plus fp64 reference of the same bits."""
    out = {}
    for cfg, nl, B, D in (('cfg1', 2, 64, None), ('cfg2', 4, 64, None), ('cfg3', 6, 32, 30), ('cfg5', 2, 32, 24)):
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            old = torch.get_default_dtype()
            torch.set_default_dtype(dtype)
            try:
                flows = cases.cfg_flow(cfg, torch.float32, n_layers=nl, D=D)   # weights/input are fp32 bits
                x = cases.cfg_input(cfg, B, torch.float32, D=D).to(dtype)
                mafs = []
                for m, sd in flows:
                    spec = m.transformer
                    if dtype == torch.float64 and isinstance(spec, fo.Spline):
                        spec = fo.Spline(x0=spec.x0.double(), xf=spec.xf.double(), n_bins=spec.n_bins,
                                         circular=spec.circular)
                    case = dict(degrees_in=m.degrees_in, spec=spec, hidden_layers=2, weight_norm=True)
                    mafs.append(to_reference_maf(ref, case, {k: v.to(dtype) for k, v in sd.items()}))
                seq = ref.SequentialFlow(*mafs)
                with torch.no_grad():
                    y, ld = seq(x)
                    out[f'{cfg}/{tag}/y'], out[f'{cfg}/{tag}/ld'] = _np(y), _np(ld)
                    if cfg != 'cfg3':
                        xi, ldi = seq.inverse(y)
                        out[f'{cfg}/{tag}/xinv'], out[f'{cfg}/{tag}/ldinv'] = _np(xi), _np(ldi)
                if tag == 'f32':
                    out[f'{cfg}/checksum'] = np.float64(sum(cases.checksum(sd) for _, sd in flows))
                    out[f'{cfg}/x'] = _np(x)
            finally:
                torch.set_default_dtype(old)
    np.savez_compressed(os.path.join(OUT, 'cfg_slices.npz'), **out)


def golden_analysis(ref):
    out = {}
    w = cases.normal((20000,), 3)
    out['w_seed3_n20000/fep'] = _np(ref.fep_estimator(w))
    out['w_seed3_n20000/fep_kT2.5'] = _np(ref.fep_estimator(w, kT=2.5))
    out['w_seed3_n20000/fep_f64'] = _np(ref.fep_estimator(w.double()))
    wb = torch.stack([w, cases.normal((20000,), 4) * 0.3], dim=1)
    out['biased/fep'] = _np(ref.fep_estimator(wb))
    for seed in (0, 1, 12345):
        g = torch.Generator().manual_seed(seed)
        out[f'randint/seed{seed}_high20000'] = _np(torch.randint(0, 20000, (3, 700), generator=g))
        out[f'randint/seed{seed}_high1e8_cont'] = _np(torch.randint(0, 100000000, (2, 700), generator=g))
    # per-resample statistics and summary of the reference for a fixed seed (batch must not matter)
    stats = torch.empty(40)
    ref.bootstrap.__globals__['_bootstrap_statistics'](w.expand(7, -1), ref.fep_estimator, 40, 20000, False,
                                                        torch.Generator().manual_seed(1), stats)
    out['bootstrap/stats_seed1_r40'] = _np(stats)
    for method in ('percentile', 'basic'):
        stat = ref.fep_estimator if method == 'percentile' else (lambda d, vectorized=False: d.mean(dim=-1))
        r = ref.bootstrap(w, stat, n_resamples=40, batch=7, method=method,
                          generator=torch.Generator().manual_seed(1))
        out[f'bootstrap/{method}'] = np.array([float(r['confidence_interval']['low']),
                                               float(r['confidence_interval']['high']),
                                               float(r['standard_deviation']), float(r['mean']), float(r['median'])])
    r = ref.bootstrap(w, ref.fep_estimator, n_resamples=30, bootstrap_sample_size=[100, 5000], take_first_only=True,
                      generator=torch.Generator().manual_seed(2))
    out['bootstrap/sizes_take_first'] = np.array([[float(x['confidence_interval']['low']),
                                                   float(x['confidence_interval']['high']),
                                                   float(x['standard_deviation']), float(x['mean']),
                                                   float(x['median'])] for x in r])
    u, ld, lw = cases.normal((64,), 5), cases.normal((64,), 6), cases.normal((64,), 7)
    out['loss/mean'] = _np(ref.BoltzmannKLDivLoss()(u, ld))
    out['loss/weighted'] = _np(ref.BoltzmannKLDivLoss()(u, ld, log_weights=lw, ref_potentials=u * 0.5))
    # every argument combination, NaN-ignoring variants and reference autograd gradients (loss.py:125-140)
    un = u.clone()
    un[[3, 17]] = float('nan')
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for nan in (0, 1):
            L = ref.BoltzmannKLDivLoss(ignore_nan=bool(nan))
            for tag, ub in (('clean', u), ('nan', un)):
                for weighted in (0, 1):
                    leaves = [t.clone().requires_grad_(True) for t in (ub, ld, lw, u * 0.5)]
                    val = L(leaves[0], leaves[1], log_weights=leaves[2] if weighted else None, ref_potentials=leaves[3])
                    key = f'loss/{tag}_w{weighted}_ignore{nan}'
                    out[key] = _np(val)
                    if torch.isfinite(val):
                        val.backward()
                        for name, t in zip(('target', 'logdet', 'logw', 'ref'), leaves):
                            if t.grad is not None:
                                out[f'{key}/grad_{name}'] = _np(torch.nan_to_num(t.grad))
    np.savez_compressed(os.path.join(OUT, 'analysis.npz'), **out)


def golden_degrees(ref):
    tab = {'generate_degrees': [], 'hidden_degrees': []}
    for n, kw in [(3, {}), (2, dict(order='descending')), (5, dict(max_value=1)),
                  (5, dict(order='descending', max_value=1)), (6, dict(conditioning_indices=[0, 3])),
                  (5, dict(order='descending', conditioning_indices=[4])), (5, dict(max_value=2, conditioning_indices=[1])),
                  (6, dict(order='descending', max_value=2, conditioning_indices=[0, 5])),
                  (7, dict(max_value=1, conditioning_indices=[0, 4], repeats=2)),
                  (6, dict(order='descending', conditioning_indices=[1, 5], repeats=3)),
                  (7, dict(conditioning_indices=[1, 2], repeats=[1, 2, 3])),
                  (6, dict(order='descending', max_value=1, conditioning_indices=[2], repeats=[1, 2])),
                  (12, dict(repeats=3, order='descending'))]:
        tab['generate_degrees'].append(dict(n_features=n, kwargs=kw, expected=ref.generate_degrees(n, **kw).tolist()))
    for din, dout, hl in [([0, 1, 2], [0, 1, 2], 1), ([0, -1, 1, 2], [0, 1, 2, 3], 2),
                          ([3, 2, 1, -1, 0], [0, 0, 1, 1, 2, 2, 3, 3], 1), ([2, -1, 0, 1], [1, 2, 0, 3] * 3, 1),
                          ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [6]), ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [6, 4]),
                          ([2, -1, 3, 0, 1], [1, 2, 0, 3] * 3, [[1, 0, -1, 2]])]:
        made = ref.MADE(torch.tensor(din), torch.tensor(dout), hl)
        tab['hidden_degrees'].append(dict(
            degrees_in=din, degrees_out=dout, hidden_layers=hl,
            expected=[h.tolist() for h in ref.MADE._get_degrees_hidden(torch.tensor(din), torch.tensor(dout), hl)],
            mask_sums=[int(made.layers[i].mask.sum()) for i in range(0, len(made.layers), 2)],
            n_parameters=int(made.n_parameters())))
    # shapes of the BASELINE.json configurations
    shapes = {}
    for name, D, t, rep in (('cfg1', 66, ref.AffineTransformer(), 1),
                            ('cfg2', 66, ref.NeuralSplineTransformer(torch.zeros(66), torch.ones(66), 8, circular=True), 1),
                            ('cfg3_sos', 300, ref.SOSPolynomialTransformer(2), 1),
                            ('cfg3_moebius', 300, ref.MoebiusTransformer(3), 3)):
        maf = ref.MAF(ref.generate_degrees(D, repeats=rep), transformer=t)
        lin = [l for l in maf._conditioner.layers if hasattr(l, 'mask')]
        shapes[name] = dict(dims=[lin[0].in_features] + [l.out_features for l in lin],
                            nnz=[int(l.mask.sum()) for l in lin])
    tab['config_shapes'] = shapes
    with open(os.path.join(OUT, 'degrees.json'), 'w') as f:
        json.dump(tab, f, indent=1)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'analysis':      # only tests/golden/analysis.npz
        golden_analysis(ref)
        return
    golden_degrees(ref)
    old = torch.get_default_dtype()
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        torch.set_default_dtype(dtype)
        try:
            golden_transformers(ref, dtype, tag)
            golden_mafs(ref, dtype, tag)
            golden_wrappers(ref, dtype, tag)
        finally:
            torch.set_default_dtype(old)
    golden_embeddings(ref)
    golden_logger(ref)
    golden_cfg(ref)
    golden_analysis(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
