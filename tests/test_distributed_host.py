"""Host logic of the N-rank path on CPU with the gloo backend (world_size 2): the collectives that combine
per-rank (max, sum) pairs and per-resample bootstrap sums.  Local partials come from the oracle here (no GPU);
on the GPU box the same helpers are fed by the kernels (tests/test_gpu_analysis.py)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partial(v):
    """(max, sum exp(v - max)) in double: what tfepb_lse returns for a shard."""
    v = v.double()
    m = v.max()
    return torch.stack([m, torch.exp(v - m).sum()])


def _worker(rank, world, port, n, kT, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import analysis_oracle as ao
        from oracle import cases
        from tfep_b200.analysis import distributed as D
        from tfep_b200.analysis.estimator import _log_n
        w = cases.normal((n,), 3) * 2.0
        lo, hi = rank * n // world, (rank + 1) * n // world
        shard = w[lo:hi]
        # estimator
        m, s = D.combine_lse_partials(_partial(-shard / kT))
        df = (-kT * (m + torch.log(s) - _log_n(n))).float()
        ref = ao.fep_estimator(w, kT=kT)
        # bootstrap: every rank walks the same global index stream, sums only its shard
        R = 12
        idx = torch.from_numpy(ao.resample_indices(7, R, n, n))
        local_max = (-shard / kT).double().max()
        e = torch.exp(-shard.double() / kT - local_max)
        inside = (idx >= lo) & (idx < hi)
        sums = torch.where(inside, e[(idx - lo).clamp(0, hi - lo - 1)], torch.zeros((), dtype=torch.double)).sum(dim=1)
        gs, gmax = D.combine_bootstrap_sums(sums, local_max)
        stats = (-kT * (gmax + torch.log(gs) - _log_n(n))).float()
        ref_stats = ao.bootstrap_statistics(w, lambda d, vectorized=False: ao.fep_estimator(d, kT=kT, vectorized=vectorized),
                                            R, generator=torch.Generator().manual_seed(7))
        # stratified Philox bootstrap: every rank derives the same cells / owners / multinomial counts
        import tfep_b200.analysis.bootstrap  # noqa: F401
        bmod = sys.modules['tfep_b200.analysis.bootstrap']
        bmod.L2_TILE_ENTRIES = 700
        cells, owner = D.shard_cells(lo, hi - lo, n, 'cpu')
        counts = bmod.stratified_counts(R, n, cells, n, seed=99)
        ok = (cells[0][0] == 0 and cells[-1][1] == n and all(a[1] == b[0] for a, b in zip(cells, cells[1:]))
              and all(b - a <= 700 for a, b in cells) and sorted(set(owner)) == [0, 1]
              and all(lo <= a and b <= hi for (a, b), o in zip(cells, owner) if o == rank)
              and bool((counts.sum(axis=1) == n).all()))
        gathered = [None, None]
        dist.all_gather_object(gathered, (cells, owner, counts.tolist()))
        ok = ok and gathered[0] == gathered[1]
        out.put((rank, float((df - ref).abs()), float((stats - ref_stats).abs().max()), ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('kT', [1.0, 2.5])
def test_two_rank_estimator_and_bootstrap_match_single_process(kT):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5001, kT, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_df, err_boot, cells_ok in res:
        assert err_df < 2e-6 and err_boot < 2e-6, (rank, err_df, err_boot)
        assert cells_ok, rank


def test_combine_partials_is_order_and_split_invariant():
    from tfep_b200.analysis.estimator import combine_partials
    v = torch.randn(1000, dtype=torch.double) * 30
    whole = _partial(v)
    for cuts in ([500], [1, 999], [10, 20, 700]):
        parts = torch.stack([_partial(c) for c in torch.tensor_split(v, cuts)])
        m, s = combine_partials(parts)
        assert abs(float(m + torch.log(s)) - float(whole[0] + torch.log(whole[1]))) < 1e-12


def _grad_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tfep_b200.utils.data_parallel import allreduce_gradients, broadcast_parameters, shard_bounds
        torch.manual_seed(100 + rank)                      # different initial weights per rank on purpose
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 3))
        broadcast_parameters(net)
        torch.manual_seed(1)
        x = torch.randn(40, 5)
        lo, hi = shard_bounds(40)
        # mean loss over the GLOBAL batch = sum over ranks of (local sum / n_global)
        loss = (net(x[lo:hi]) ** 2).sum() / 40
        loss.backward()
        n = allreduce_gradients(net, average=False)
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 3))
        ref.load_state_dict(net.state_dict())
        ((ref(x) ** 2).sum() / 40).backward()
        err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        out.put((rank, err, n, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    """Data-parallel training step: gradients of the batch-sharded loss, added with one flat all-reduce,
    equal the single-process gradients of the full batch (weights broadcast from rank 0 first)."""
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[3] for r in res) == [(0, 20), (20, 40)]
    for rank, err, n, _ in res:
        assert err < 1e-6 and n == 5 * 7 + 7 + 7 * 3 + 3, (rank, err, n)
