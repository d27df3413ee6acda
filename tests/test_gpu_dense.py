"""GPU parity: DMMA GEMM, blocked Cholesky + inverse, LAUUM vs NumPy/LAPACK (fp64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eng():
    from portfoliooptgp_b200 import ops
    return ops.shared_engine(0)


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (130, 70, 33), (257, 384, 129), (1000, 900, 515)])
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm(M, N, K, ta, tb):
    import torch
    rng = np.random.default_rng(0)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    dA, dB, dC = (torch.tensor(x, device="cuda") for x in (A, B, C0))
    eng = _eng()
    eng.gemm(ta, tb, M, N, K, 0.7, dA.data_ptr(), dA.shape[1], dB.data_ptr(), dB.shape[1], -0.3, dC.data_ptr(), N)
    torch.cuda.synchronize()
    want = 0.7 * ((A.T if ta else A) @ (B.T if tb else B)) - 0.3 * C0
    err = np.max(np.abs(dC.cpu().numpy() - want))
    assert err < 1e-11 * max(1, K), err


def test_gemm_large_tiles():
    import torch
    rng = np.random.default_rng(1)
    M = N = 2048; K = 512
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    dA, dB = torch.tensor(A, device="cuda"), torch.tensor(B, device="cuda")
    dC = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    _eng().gemm(0, 1, M, N, K, 1.0, dA.data_ptr(), K, dB.data_ptr(), K, 0.0, dC.data_ptr(), N)
    torch.cuda.synchronize()
    assert np.max(np.abs(dC.cpu().numpy() - A @ B.T)) < 1e-10


# every leaf shape 1..128 rows that changes the tile grid of the in-shared-memory factorisation (1 to 16
# panels, odd and even super-tile counts, ragged edges) plus multi-leaf sizes
@pytest.mark.parametrize("n", [1, 2, 5, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 40, 47, 48, 56, 63, 64, 65, 72, 80, 89, 96,
                               104, 111, 112, 120, 121, 127, 128, 129, 200, 255, 256, 257, 300, 1000, 2500])
def test_potrf_inv_lauum(n):
    import torch
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n + 7))
    S = G @ G.T / n + 0.05 * np.eye(n)
    ld = (n + 15) // 16 * 16
    dA = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    dA[:, :n] = torch.tensor(np.tril(S), device="cuda")
    dW = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    dO = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    eng = _eng()
    eng.potrf_inv(dA.data_ptr(), n, ld, dW.data_ptr(), ld)
    eng.lauum(dW.data_ptr(), n, ld, dO.data_ptr(), ld)
    torch.cuda.synchronize()
    L = np.tril(dA.cpu().numpy()[:, :n])
    W = np.tril(dW.cpu().numpy()[:, :n])
    Kinv = np.tril(dO.cpu().numpy()[:, :n])
    Lref = np.linalg.cholesky(S)
    scale = np.linalg.cond(S)
    assert np.max(np.abs(L - Lref)) < 1e-14 * scale * 10
    assert np.max(np.abs(W @ Lref - np.eye(n))) < 1e-14 * scale * 10
    assert np.max(np.abs(Kinv - np.tril(np.linalg.inv(S)))) < 1e-14 * scale * np.max(np.abs(np.linalg.inv(S))) * 10


def test_potrf_reports_non_pd():
    import torch
    import portfoliooptgp_b200 as gpflow
    n = 200
    S = np.eye(n); S[150, 150] = -1.0
    dA = torch.tensor(S, device="cuda")
    with pytest.raises(gpflow.CholeskyError):
        _eng().potrf(dA.data_ptr(), n, n)


# the factor-only path (N^3/3 flop, gpb_potrf): one diagonal block, exactly two, ragged multi-block sizes
@pytest.mark.parametrize("n", [5, 128, 1000, 1024, 1025, 2048, 2500, 3100, 5000])
def test_potrf_factor_only(n):
    import torch
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n + 7))
    S = G @ G.T / n + 0.05 * np.eye(n)
    ld = (n + 15) // 16 * 16
    dA = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    dA[:, :n] = torch.tensor(np.tril(S), device="cuda")
    _eng().potrf(dA.data_ptr(), n, ld)
    torch.cuda.synchronize()
    L = np.tril(dA.cpu().numpy()[:, :n])
    Lref = np.linalg.cholesky(S)
    assert np.max(np.abs(L - Lref)) < 1e-14 * np.linalg.cond(S) * 10


def test_potrf_factor_only_reports_pivot_row_in_a_later_block():
    import torch
    import portfoliooptgp_b200 as gpflow
    n = 2300
    S = np.eye(n); S[2100, 2100] = -1.0
    dA = torch.tensor(S, device="cuda")
    with pytest.raises(gpflow.CholeskyError) as ei:
        _eng().potrf(dA.data_ptr(), n, n)
    assert "2101" in str(ei.value)
