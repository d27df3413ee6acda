"""CPU: lock-step L-BFGS-B reproduces SciPy's iterates; sharding + gather over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import scipy.optimize

from portfoliooptgp_b200.batched import gather_results, lockstep_lbfgsb, shard_range


def _rosen_family(a):
    def f(x):
        return float(np.sum(a * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))

    def g(x):
        out = np.zeros_like(x)
        out[:-1] += -4 * a * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        out[1:] += 2 * a * (x[1:] - x[:-1] ** 2)
        return out
    return f, g


def test_lockstep_matches_scipy_iterates():
    rng = np.random.default_rng(0)
    B, n = 7, 5
    coeffs = rng.uniform(1.0, 100.0, size=B)
    X0 = rng.uniform(-1.5, 1.5, size=(B, n))
    funs = [_rosen_family(a) for a in coeffs]
    calls = []

    def fun_batch(X, idx):
        calls.append(len(idx))
        f = np.array([funs[b][0](x) for x, b in zip(X, idx)])
        g = np.stack([funs[b][1](x) for x, b in zip(X, idx)])
        return f, g

    res = lockstep_lbfgsb(fun_batch, X0, maxiter=60)
    for b in range(B):
        ref = scipy.optimize.minimize(lambda x: (funs[b][0](x), funs[b][1](x)), X0[b], jac=True, method="L-BFGS-B",
                                      options=dict(maxiter=60))
        assert res[b].nit == ref.nit and res[b].nfev == ref.nfev
        assert np.array_equal(res[b].x, ref.x)          # bit-identical iterates
        assert res[b].fun == ref.fun and res[b].status == ref.status
    assert calls[0] == B and min(calls) >= 1 and len(calls) < sum(r.nfev for r in res)  # batched rounds


def test_shard_range_partitions():
    for total, world in [(5120, 8), (5120, 3), (7, 8), (10, 4)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(5120, 3, 8) == (1920, 2560)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 11
    lo, hi = shard_range(total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]], dtype=torch.float64)
    full = gather_results(local, total)
    q.put((rank, full.numpy()))
    dist.destroy_process_group()


def test_gather_results_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(11, dtype=np.float64)[:, None] * np.array([[1.0, 10.0, 100.0]])
    for _, full in outs:
        assert np.array_equal(full, want)


def test_lockstep_workers_reproduce_the_in_process_iterates():
    """workers > 1: slices of the batch advance in worker processes (_lbfgsb_pool); every result field is
    bit-identical to the in-process driver (and therefore to scipy.optimize.minimize)."""
    from portfoliooptgp_b200 import _lbfgsb_pool
    from portfoliooptgp_b200.batched import lockstep_lbfgsb
    B, n = 64, 4
    rng = np.random.default_rng(5)
    A = rng.uniform(0.5, 2.0, size=(B, n))
    c = rng.normal(size=(B, n))

    def fun(U, idx):
        d = U - c[idx]
        return 0.5 * np.sum(A[idx] * d * d, axis=1) + 0.1 * np.sum(d ** 4, axis=1) + np.sum(np.cos(d), axis=1), \
            A[idx] * d + 0.4 * d ** 3 - np.sin(d)

    in_flight = []

    def fun_async(U, idx):
        U, idx = U.copy(), idx.copy()
        in_flight.append(1)
        assert len(in_flight) <= 2

        def wait():
            in_flight.pop()
            return fun(U, idx)
        return wait

    X0 = rng.normal(size=(B, n))
    r0 = lockstep_lbfgsb(fun, X0, maxiter=60)
    try:
        r1 = lockstep_lbfgsb(fun, X0, maxiter=60, workers=3)
        # two groups of workers alternating: one group's evaluation in flight while the other group advances
        r2 = lockstep_lbfgsb(fun, X0, maxiter=60, workers=4, fun_batch_async=fun_async)
    finally:
        _lbfgsb_pool.shutdown()
    assert len(r1) == B and len(r2) == B
    for r in (r1, r2):
        for a, b in zip(r0, r):
            assert np.array_equal(a.x, b.x) and a.fun == b.fun and np.array_equal(a.jac, b.jac)
            assert (a.nit, a.nfev, a.status, a.message) == (b.nit, b.nfev, b.status, b.message)


def test_pipelined_halves_give_the_same_iterates():
    """Two half-batches in flight (device evaluates one while the host advances the other): every problem still
    sees its own (x, f, g) sequence, so the results equal the single-batch lock-step run bit for bit."""
    rng = np.random.default_rng(3)
    B, n = 11, 4
    coeffs = rng.uniform(1.0, 100.0, size=B)
    X0 = rng.uniform(-1.5, 1.5, size=(B, n))
    funs = [_rosen_family(a) for a in coeffs]
    in_flight = []

    def fun_batch(X, idx):
        return (np.array([funs[b][0](x) for x, b in zip(X, idx)]), np.stack([funs[b][1](x) for x, b in zip(X, idx)]))

    def fun_batch_async(X, idx):
        X, idx = X.copy(), idx.copy()
        in_flight.append(1)

        def wait():
            in_flight.pop()
            return fun_batch(X, idx)
        assert len(in_flight) <= 2
        return wait

    a = lockstep_lbfgsb(fun_batch, X0, maxiter=60)
    b = lockstep_lbfgsb(fun_batch, X0, maxiter=60, fun_batch_async=fun_batch_async)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.x, rb.x) and ra.fun == rb.fun and ra.nit == rb.nit and ra.nfev == rb.nfev and ra.status == rb.status


def test_old_scipy_gets_a_clear_error(monkeypatch):
    """the driver uses SciPy's private C setulb (>= 1.15); anything else must fail with a message,
    not with a TypeError from inside the first fit."""
    import pytest
    from portfoliooptgp_b200 import batched
    monkeypatch.setattr(batched, "_SCIPY_CHECKED", None)
    monkeypatch.setattr(batched.scipy, "__version__", "1.13.0")
    with pytest.raises(RuntimeError, match="SciPy >= 1.15"):
        lockstep_lbfgsb(lambda X, idx: (np.zeros(len(idx)), np.zeros_like(X)), np.zeros((2, 2)))
    monkeypatch.setattr(batched, "_SCIPY_CHECKED", None)
