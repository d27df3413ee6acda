"""Device-side data preparation (csrc/prep.cu, data_prep.py) against the pandas restatement of
Multi-Input_GPR/utils/data_handler.py in oracle/data_prep_oracle.py."""
import numpy as np
import pytest

from oracle import data_prep_oracle as O


def _prices(T, A, seed):
    rng = np.random.default_rng(seed)
    close = 100.0 * np.exp(np.cumsum(0.01 * rng.normal(size=(T, A)), axis=0))
    open_ = close * (1 + 0.003 * rng.normal(size=(T, A)))
    return close, open_


# ---- oracle known answers (CPU) ---------------------------------------------------------------
def test_oracle_known_answers():
    close = np.array([100.0, 110.0, 99.0, 99.0])
    r = O.returns(close, kind="return")
    np.testing.assert_allclose(r, [0.1, 0.1, -0.1, 0.0], atol=1e-15)      # row 0 <- row 1 (:87-88)
    lr = O.returns(close, kind="daily_log_return")
    assert np.isnan(lr[0])
    np.testing.assert_allclose(lr[1:], np.log([1.1, 0.9, 1.0]), atol=1e-15)
    z, m, s = O.zscore(np.array([1.0, 2.0, 3.0, 4.0]))
    assert m[0] == 2.5 and abs(s[0] - np.sqrt(5.0 / 3.0)) < 1e-15           # ddof = 1
    X, Y = O.rolling_windows(np.arange(12.0).reshape(6, 2), np.arange(6.0), 4, 1)
    assert X.shape == (3, 4, 2) and Y.shape == (3, 4, 1) and X[2, 0, 0] == 4.0 and Y[1, 3, 0] == 4.0


def test_concatenate_errors_without_gpu():
    from portfoliooptgp_b200 import data_prep
    with pytest.raises(ValueError, match="list or tuple"):
        data_prep.concatenate_X(np.zeros((3, 1)))
    with pytest.raises(ValueError, match="at least one"):
        data_prep.concatenate_X([])
    with pytest.raises(ValueError, match="kind must be"):
        data_prep.returns(np.ones(4), kind="close")


# ---- GPU parity --------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("T,A", [(2, 1), (257, 1), (1000, 8), (5000, 40), (1024, 33)])
@pytest.mark.parametrize("kind", ["return", "intraday_return", "daily_log_return"])
def test_returns_match_pandas(T, A, kind):
    from portfoliooptgp_b200 import data_prep
    close, open_ = _prices(T, A, seed=T + A)
    if kind == "daily_log_return" and T > 3:
        close[2, 0] = 0.0    # log(0 / c) = -inf and log(c / 0) = +inf -> 0 (:91)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = O.returns(close, open_, kind)
    got = data_prep.returns(close, open_, kind).cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    # x / y - 1 and log(x / y) are correctly rounded ops on both sides except libm vs CUDA log (<= 1 ulp)
    np.testing.assert_allclose(got, ref, rtol=4e-16, atol=4e-16, equal_nan=True)


@pytest.mark.gpu
def test_returns_1d_series():
    from portfoliooptgp_b200 import data_prep
    close, _ = _prices(300, 1, seed=3)
    got = data_prep.returns(close[:, 0]).cpu().numpy()
    assert got.shape == (300,)
    np.testing.assert_allclose(got, O.returns(close[:, 0]), rtol=4e-16, atol=4e-16)


@pytest.mark.gpu
@pytest.mark.parametrize("T,A", [(2, 1), (1000, 1), (1025, 8), (8192, 8), (70000, 3), (3000, 70)])
def test_zscore_matches_pandas(T, A):
    from portfoliooptgp_b200 import data_prep
    rng = np.random.default_rng(T * 7 + A)
    x = 5.0 + 3.0 * rng.normal(size=(T, A)) + np.arange(A)
    z_ref, m_ref, s_ref = O.zscore(x)
    z, m, s = data_prep.zscore(x)
    # summation order differs from pandas' (fixed two-stage tree here): a few ulp of the sums
    np.testing.assert_allclose(m.cpu().numpy(), m_ref, rtol=1e-13)
    np.testing.assert_allclose(s.cpu().numpy(), s_ref, rtol=1e-13)
    np.testing.assert_allclose(z.cpu().numpy(), z_ref, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_zscore_is_deterministic_and_ddof0():
    from portfoliooptgp_b200 import data_prep
    x = np.random.default_rng(0).normal(size=(4099, 5))
    z1, m1, s1 = data_prep.zscore(x)
    z2, m2, s2 = data_prep.zscore(x)
    assert (z1 == z2).all() and (m1 == m2).all() and (s1 == s2).all()
    _, _, s0 = data_prep.zscore(x, ddof=0)
    np.testing.assert_allclose(s0.cpu().numpy(), x.std(axis=0), rtol=1e-13)


@pytest.mark.gpu
def test_reference_pipeline_design_matrix():
    """process_data -> normalize_and_reshape -> concatenate_X as Multi-Input_GPR/main.py:380-410 does
    for several tickers, fused into one [T, D] matrix on the device."""
    from portfoliooptgp_b200 import data_prep
    T = 700
    close, open_ = _prices(T, 3, seed=9)
    day = np.arange(T, dtype=np.float64) * 1.4
    # reference order: per ticker, z-score the return column; then the time column
    cols_ref = [O.zscore(O.returns(close[:, a]))[0] for a in range(3)] + [O.zscore(day)[0]]
    X_ref = O.concatenate_X(cols_ref)
    r = data_prep.returns(close)
    X, mean, std = data_prep.design_matrix([r, day])
    assert X.shape == (T, 4) and X.is_contiguous()
    np.testing.assert_allclose(X.cpu().numpy(), X_ref, rtol=1e-12, atol=1e-12)
    # the unfused route gives the same bits
    parts = [data_prep.zscore(r[:, a].contiguous())[0] for a in range(3)] + [data_prep.zscore(day)[0]]
    X2 = data_prep.concatenate_X(parts)
    assert (X2 == X).all()
    Xn, Yn, (ym, ys), (xm, xs) = data_prep.normalize_and_reshape(r[:, 0].contiguous(), day)
    assert Xn.shape == (T, 1) and Yn.shape == (T, 1)
    assert (Yn[:, 0] == X[:, 0]).all() and (Xn[:, 0] == X[:, 3]).all()
    with pytest.raises(ValueError, match="same shape"):
        data_prep.concatenate_X([parts[0], parts[1][:-1]])


@pytest.mark.gpu
@pytest.mark.parametrize("S,T,D,N,stride", [(1, 191, 8, 128, 1), (20, 191, 8, 128, 1), (3, 100, 1, 7, 5),
                                            (2, 64, 16, 64, 3), (2, 10, 2, 11, 1)])
def test_rolling_windows_bit_exact(S, T, D, N, stride):
    from portfoliooptgp_b200 import data_prep
    rng = np.random.default_rng(S + T + D)
    f = rng.normal(size=(S, T, D))
    y = rng.normal(size=(S, T))
    X_ref, Y_ref = O.rolling_windows(f, y, N, stride)
    X, Y = data_prep.rolling_windows(f, y, window=N, stride=stride)
    assert X.shape == X_ref.shape and Y.shape == Y_ref.shape
    assert np.array_equal(X.cpu().numpy(), X_ref) and np.array_equal(Y.cpu().numpy(), Y_ref)
    X_only = data_prep.rolling_windows(f, window=N, stride=stride)
    assert np.array_equal(X_only.cpu().numpy(), X_ref)


@pytest.mark.gpu
def test_windows_feed_batched_gpr():
    """Device-built [B, N, D] batch == host-built batch through BatchedGPR (C3 data path)."""
    import portfoliooptgp_b200 as gpflow
    from portfoliooptgp_b200 import data_prep
    close, _ = _prices(160, 4, seed=21)
    r = data_prep.returns(close)
    feat, _, _ = data_prep.design_matrix([r[:, 1:].contiguous(), np.arange(160.0)])
    yz, _, _ = data_prep.zscore(r[:, 0].contiguous())
    X, Y = data_prep.rolling_windows(feat, yz, window=128, stride=4)
    B = X.shape[0]
    assert B == 9
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52()
    f_dev = gpflow.BatchedGPR(X, Y, k, noise_variance=0.1).lml_and_grads()
    Xh, Yh = O.rolling_windows(feat.cpu().numpy(), yz.cpu().numpy(), 128, 4)
    f_host = gpflow.BatchedGPR(Xh, Yh, k, noise_variance=0.1).lml_and_grads()
    for a, b in zip(f_dev, f_host):
        assert np.array_equal(a, b)
    assert np.isfinite(f_dev[0]).all() and (f_dev[3] == 0).all()
