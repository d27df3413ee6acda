"""CPU, gloo world_size 2: the data-parallel SVGP reduction (sum the unscaled data-term records,
scale once by num_data/(world*B), subtract the KL once) reproduces the single-process ELBO."""
import os
import socket

import numpy as np

from oracle import gpflow_oracle as O


def _problem():
    rng = np.random.default_rng(0)
    N, M = 64, 9
    X = rng.standard_normal((N, 2)); Y = rng.standard_normal((N, 1))
    Z = X[:M] + 0.05
    qmu = 0.2 * rng.standard_normal((M, 1))
    qs = (0.7 * np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))))[None]
    k = O.Sum([O.Leaf("se", 1.1, 0.9), O.Leaf("linear", 0.2)])
    return k, X, Y, Z, qmu, qs


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from portfoliooptgp_b200.svgp_dp import allreduce_sum_, combine_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, X, Y, Z, qmu, qs = _problem()
    B = X.shape[0] // world
    Xr, Yr = X[rank * B:(rank + 1) * B], Y[rank * B:(rank + 1) * B]
    kl = O.gauss_kl(qmu, qs)
    S_local = O.svgp_elbo(k, Z, qmu, qs, 0.1, Xr, Yr, num_data=None) + kl      # unscaled data term of this rank
    flat = torch.tensor([S_local, float(rank + 1)], dtype=torch.float64)
    allreduce_sum_(flat)
    elbo = combine_records(float(flat[0]), kl, num_data=200, world=world, B=B)
    q.put((rank, elbo, float(flat[1])))
    dist.destroy_process_group()


def test_data_parallel_elbo_reduction_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, X, Y, Z, qmu, qs = _problem()
    want = O.svgp_elbo(k, Z, qmu, qs, 0.1, X, Y, num_data=200)   # one process, the union of both minibatches
    for _, elbo, tag in outs:
        assert abs(elbo - want) <= 1e-12 * abs(want)
        assert tag == 3.0


def test_shard_windows_visit_every_row_and_keep_rows_paired():
    """the window cursor used to reset at the end of the shard, so the last n mod B rows were never
    drawn.  With the per-epoch permutation every row is drawn equally often, and X rows stay with their targets."""
    import torch
    from portfoliooptgp_b200.svgp_dp import ShardWindows
    n, B = 10, 4
    X = torch.arange(n, dtype=torch.float64)[:, None].repeat(1, 3)
    y = 100.0 + torch.arange(n, dtype=torch.float64)
    w = ShardWindows(X, y, B, seed=1)
    counts = np.zeros(n)
    for _ in range(4000):
        o = w.next()
        xb, yb = X[o:o + B], y[o:o + B]
        assert torch.equal(xb[:, 0] + 100.0, yb) and torch.equal(xb[:, 0], xb[:, 2])
        counts[xb[:, 0].long().numpy()] += 1
    assert w.epoch > 1000
    assert counts.min() > 0.85 * counts.mean() and counts.max() < 1.15 * counts.mean()
    # without shuffling the old behaviour (fixed windows) is still available, explicitly
    w2 = ShardWindows(X.clone(), y.clone(), B, shuffle=False)
    assert [w2.next() for _ in range(4)] == [0, 4, 0, 4]


def _worker_unequal(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from portfoliooptgp_b200.svgp_dp import check_equal_minibatch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        check_equal_minibatch(64, torch.device("cpu"))
        ok_same = True
        try:
            check_equal_minibatch(64 + rank, torch.device("cpu"))
            raised = False
        except ValueError:
            raised = True
        q.put((rank, ok_same, raised))
    finally:
        dist.destroy_process_group()


def test_unequal_minibatch_sizes_are_rejected_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_unequal, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and raised for _, ok, raised in outs)


def _worker_packed(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from portfoliooptgp_b200.svgp_dp import PackedRecord, allreduce_sum_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, D, P = 13, 3, 4
        head = 2 + P + M * D + M
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(head + M * M, dtype=torch.float64, generator=g)
        want = flat.clone()
        allreduce_sum_(want)                                        # the plain, full-size reduction
        pr = PackedRecord(M, D, P, torch.device("cpu"))
        got = pr.allreduce_(flat.clone())
        low = torch.tril(torch.ones(M, M, dtype=torch.bool)).reshape(-1)
        ok_head = torch.equal(got[:head], want[:head])
        ok_low = torch.equal(got[head:][low], want[head:][low])
        untouched = torch.equal(got[head:][~low], flat[head:][~low])   # the strict upper part is not sent
        q.put((rank, ok_head, ok_low, untouched, pr.buf.numel(), head + M * (M + 1) // 2))
    finally:
        dist.destroy_process_group()


def test_packed_record_allreduce_equals_full_allreduce_on_the_meaningful_part_gloo_world2():
    """the q_sqrt gradient crosses the links as its lower triangle (M (M + 1) / 2 doubles)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, ok_head, ok_low, untouched, n_sent, n_want in outs:
        assert ok_head and ok_low and untouched
        assert n_sent == n_want
