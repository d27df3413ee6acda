// prep.cu -- device-side data preparation: the step immediately before the GP hot path
// (SURVEY.md section 8f rank 4).  Price series that are already in HBM are turned into returns,
// z-scored per column, joined into the [T, D] design matrix and cut into the [B, N, D] window batch
// that gpb_batched_lml_grad consumes, without a host round trip.
//
// Replaces the pandas arithmetic of Multi-Input_GPR/utils/data_handler.py:
//   :86-88  df['close'].pct_change(), first row filled with the second row's value   (kind 0)
//   :89     (close - open) / open                                                    (kind 1)
//   :90-91  log(close / close.shift(1)), +-inf -> 0, first row NaN                   (kind 2)
//   :160-169 normalize_and_reshape: (x - mean) / std with pandas' ddof = 1 std
//   :129-154 concatenate_X: column concat of [T,1] series
// and the window slicing of Multi-Input_GPR/main.py:414-423 / the C3 configuration (stride-s windows
// of length N over each asset's series).
//
// All kernels are HBM-bound byte movers with a handful of FP64 ops per element: flat, fully
// coalesced indexing over row-major [T, A] arrays; column statistics are deterministic two-stage
// reductions (fixed partition, fixed order), so results do not depend on the launch.
#include <math.h>

#include "engine.cuh"

namespace gpb {

// ------------------------------------------------------------------------------------------------
// returns: series are [T, A] row-major, one column per asset / field
template <int KIND>
__global__ void returns_kernel(const double* __restrict__ close, const double* __restrict__ open, int64_t T, int64_t A,
                               double* __restrict__ out) {
    const int64_t total = T * A;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double c = close[i];
        double r;
        if (KIND == 1) {
            const double o = open[i];
            r = (c - o) / o;
        } else {
            const int64_t t = i / A;
            if (KIND == 0) {
                // pct_change; row 0 takes row 1's value (fillna with return.iloc[1])
                if (T < 2) r = nan("");
                else if (t == 0) r = close[i + A] / c - 1.0;
                else r = c / close[i - A] - 1.0;
            } else {
                if (t == 0) {
                    r = nan("");
                } else {
                    r = log(c / close[i - A]);
                    if (isinf(r)) r = 0.0;
                }
            }
        }
        out[i] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// column statistics of a [T, A] array.  Stage 1: CTA (cx, cy) sums rows [cy*RPB, (cy+1)*RPB) of
// columns [cx*32, cx*32+32) -- 32 x 8 threads, lanes along the columns (coalesced), a fixed-order
// shared-memory fold over the 8 row groups.  Stage 2: one thread per column folds the partials in
// order.  PASS 0 sums x, PASS 1 sums (x - mean)^2.
constexpr int PREP_RPB = 1024;

template <int PASS>
__global__ void colstat_partial_kernel(const double* __restrict__ x, int64_t T, int64_t A, const double* __restrict__ mean,
                                       double* __restrict__ part) {
    __shared__ double sh[8][33];
    const int64_t col = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * PREP_RPB;
    const int64_t r1 = (r0 + PREP_RPB < T) ? r0 + PREP_RPB : T;
    double s = 0.0;
    if (col < A) {
        const double m = PASS ? mean[col] : 0.0;
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            const double v = x[r * A + col] - m;
            s += PASS ? v * v : v;
        }
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < A) {
        double t = sh[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sh[k][threadIdx.x];
        part[(int64_t)blockIdx.y * A + col] = t;
    }
}

// PASS 0: mean = sum / T.  PASS 1: std = sqrt(sum / (T - ddof)).
template <int PASS>
__global__ void colstat_final_kernel(const double* __restrict__ part, int64_t nchunks, int64_t T, int64_t A, int ddof,
                                     double* __restrict__ out) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= A) return;
    double s = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) s += part[c * A + col];
    out[col] = PASS ? sqrt(s / (double)(T - ddof)) : s / (double)T;
}

// out[t, a] = (x[t, a] - mean[a]) / std[a], written with row stride ldo at column offset (so that
// several z-scored blocks land side by side in one [T, D] design matrix: concatenate_X fused)
__global__ void zscore_apply_kernel(const double* __restrict__ x, int64_t T, int64_t A, const double* __restrict__ mean,
                                    const double* __restrict__ sd, double* __restrict__ out, int64_t ldo) {
    const int64_t total = T * A;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / A, a = i - t * A;
        out[t * ldo + a] = (x[i] - mean[a]) / sd[a];
    }
}

// ------------------------------------------------------------------------------------------------
// window gather: feat [S, T, D], y [S, T]  ->  X [S*W, N, D], Y [S*W, N]; window w of series s
// covers rows w*stride .. w*stride + N - 1.  Every window is one contiguous N*D block of feat, so
// reads and writes are both unit-stride; overlapping windows re-read through L2.
__global__ void window_gather_kernel(const double* __restrict__ feat, const double* __restrict__ y, int64_t S, int64_t T,
                                     int D, int64_t N, int64_t stride, int64_t W, double* __restrict__ X,
                                     double* __restrict__ Y) {
    const int64_t per = N * D;
    const int64_t totalX = S * W * per;
    const int64_t totalY = (y && Y) ? S * W * N : 0;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < totalX + totalY; i += step) {
        if (i < totalX) {
            const int64_t b = i / per, e = i - b * per;
            const int64_t s = b / W, w = b - s * W;
            X[i] = feat[(s * T + w * stride) * D + e];
        } else {
            const int64_t k = i - totalX;
            const int64_t b = k / N, n = k - b * N;
            const int64_t s = b / W, w = b - s * W;
            Y[k] = y[s * T + w * stride + n];
        }
    }
}

static unsigned flat_grid(gpb_handle* h, int64_t total, int threads) {
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)h->sm_count * 16;  // a few resident CTAs per SM, grid-stride beyond that
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

int prep_returns(gpb_handle* h, const double* d_close, const double* d_open, int64_t T, int64_t A, int kind,
                 double* d_out) {
    if (T <= 0 || A <= 0) return 0;
    const unsigned g = flat_grid(h, T * A, 256);
    if (kind == 0) returns_kernel<0><<<g, 256, 0, h->stream>>>(d_close, d_open, T, A, d_out);
    else if (kind == 1) returns_kernel<1><<<g, 256, 0, h->stream>>>(d_close, d_open, T, A, d_out);
    else returns_kernel<2><<<g, 256, 0, h->stream>>>(d_close, d_open, T, A, d_out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "returns_kernel launch");
}

int prep_zscore(gpb_handle* h, const double* d_x, int64_t T, int64_t A, int ddof, double* d_out, int64_t ldo,
                double* d_mean, double* d_std) {
    if (T <= 0 || A <= 0) return 0;
    const int64_t nchunks = (T + PREP_RPB - 1) / PREP_RPB;
    if (nchunks > 65535) return set_error(h, -2, "zscore: T=%lld too long (max %d rows)", (long long)T, 65535 * PREP_RPB);
    double* part = workspace(h, BUF_RED, (size_t)(nchunks * A + 2 * A) * sizeof(double));
    if (!part) return -1;
    double* mean = d_mean ? d_mean : part + nchunks * A;
    double* sd = d_std ? d_std : part + nchunks * A + A;
    const dim3 grid((unsigned)((A + 31) / 32), (unsigned)nchunks), block(32, 8);
    const unsigned gf = (unsigned)((A + 127) / 128);
    colstat_partial_kernel<0><<<grid, block, 0, h->stream>>>(d_x, T, A, nullptr, part);
    colstat_final_kernel<0><<<gf, 128, 0, h->stream>>>(part, nchunks, T, A, ddof, mean);
    colstat_partial_kernel<1><<<grid, block, 0, h->stream>>>(d_x, T, A, mean, part);
    colstat_final_kernel<1><<<gf, 128, 0, h->stream>>>(part, nchunks, T, A, ddof, sd);
    h->launches += 4;
    if (d_out) {
        zscore_apply_kernel<<<flat_grid(h, T * A, 256), 256, 0, h->stream>>>(d_x, T, A, mean, sd, d_out, ldo);
        h->launches += 1;
    }
    return check_cuda(h, cudaGetLastError(), "zscore kernels launch");
}

int prep_windows(gpb_handle* h, const double* d_feat, const double* d_y, int64_t S, int64_t T, int D, int64_t N,
                 int64_t stride, double* d_X, double* d_Y) {
    const int64_t W = (T - N) / stride + 1;
    const int64_t total = S * W * N * (D + ((d_y && d_Y) ? 1 : 0));
    window_gather_kernel<<<flat_grid(h, total, 256), 256, 0, h->stream>>>(d_feat, d_y, S, T, D, N, stride, W, d_X, d_Y);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "window_gather_kernel launch");
}

// ---- prediction post-processing (GPR/predictor.py:10-51) -----------------------------------------
// upsample_predictions: pd.Series(pred, index=X).reindex(X_daily).interpolate('linear') -- exact-value lookup of
// the sparse (weekly / monthly) points in the daily grid, then linear interpolation over POSITIONS in the
// daily grid (pandas ignores the index values), NaN before the first matched point, the last value repeated
// after the last one.  Both grids ascending and unique (dates).  One thread per daily position: binary search
// for its bracketing sparse points, skipping sparse values that are not on the daily grid; Q prediction
// columns (f_mean, f_var, y_mean, y_var) share the search.  The interpolation is numpy's
// slope * (x - x_lo) + y_lo with separately rounded operations, so the result is bit-identical to pandas.
// HBM-bound byte mover: 8 (1 + Q) bytes per daily position out/in plus the (L2-resident) sparse series.
__device__ __forceinline__ int64_t lower_bound_d(const double* __restrict__ a, int64_t n, double v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void post_upsample_kernel(const double* __restrict__ Xd, int64_t Nd, const double* __restrict__ Xs, int64_t Ns,
                                     const double* __restrict__ pred, int Q, double* __restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < Nd; j += (int64_t)gridDim.x * blockDim.x) {
        const double x = Xd[j];
        // sparse points on the grid at or before / at or after daily position j
        int64_t i_hi = lower_bound_d(Xs, Ns, x), p_hi = -1;
        int64_t i_lo = (i_hi < Ns && Xs[i_hi] == x) ? i_hi : i_hi - 1, p_lo = -1;
        for (; i_lo >= 0; --i_lo) {          // (skips sparse values that are not daily grid values)
            const int64_t p = lower_bound_d(Xd, Nd, Xs[i_lo]);
            if (p < Nd && Xd[p] == Xs[i_lo]) { p_lo = p; break; }
        }
        for (; i_hi < Ns; ++i_hi) {
            const int64_t p = lower_bound_d(Xd, Nd, Xs[i_hi]);
            if (p < Nd && Xd[p] == Xs[i_hi]) { p_hi = p; break; }
        }
        for (int q = 0; q < Q; ++q) {
            const double* pq = pred + (int64_t)q * Ns;
            double v;
            if (p_lo < 0) v = nan;                               // before the first matched point
            else if (p_hi < 0 || p_hi == p_lo) v = pq[i_lo];     // after the last one, or on a matched point
            else {
                const double slope = __ddiv_rn(__dsub_rn(pq[i_hi], pq[i_lo]), (double)(p_hi - p_lo));
                v = __dadd_rn(__dmul_rn(slope, (double)(j - p_lo)), pq[i_lo]);
            }
            out[(int64_t)q * Nd + j] = v;
        }
    }
}

// out = alpha * daily + beta * weekly + (1 - alpha - beta) * monthly, in the reference's evaluation order
// (GPR/predictor.py:27-31), separately rounded: bit-identical to the NumPy expression.
__global__ void post_blend_kernel(double alpha, double beta, const double* __restrict__ d, const double* __restrict__ w,
                                  const double* __restrict__ m, int64_t n, double* __restrict__ out) {
    const double gamma = __dsub_rn(__dsub_rn(1.0, alpha), beta);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, d[i]), __dmul_rn(beta, w[i])), __dmul_rn(gamma, m[i]));
}

int post_upsample(gpb_handle* h, const double* d_Xd, int64_t Nd, const double* d_Xs, int64_t Ns, const double* d_pred, int Q,
                  double* d_out) {
    if (Nd == 0) return 0;
    post_upsample_kernel<<<flat_grid(h, Nd, 256), 256, 0, h->stream>>>(d_Xd, Nd, d_Xs, Ns, d_pred, Q, d_out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "post_upsample_kernel launch");
}

int post_blend(gpb_handle* h, double alpha, double beta, const double* d_d, const double* d_w, const double* d_m, int64_t n,
               double* d_out) {
    if (n == 0) return 0;
    post_blend_kernel<<<flat_grid(h, n, 256), 256, 0, h->stream>>>(alpha, beta, d_d, d_w, d_m, n, d_out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "post_blend_kernel launch");
}

}  // namespace gpb

extern "C" {

int gpb_post_upsample(gpb_handle* h, const double* d_Xdaily, int64_t Nd, const double* d_X, int64_t Ns, const double* d_pred,
                      int Q, double* d_out) {
    GPB_ENTER(h);
    if (Nd < 0 || Ns < 0 || Q < 1) return gpb::set_error(h, -2, "post_upsample: bad sizes");
    if (Nd > 0 && (!d_Xdaily || !d_out)) return gpb::set_error(h, -2, "post_upsample: null pointer");
    if (Ns > 0 && (!d_X || !d_pred)) return gpb::set_error(h, -2, "post_upsample: null pointer");
    return gpb::post_upsample(h, d_Xdaily, Nd, d_X, Ns, d_pred, Q, d_out);
}

int gpb_post_blend(gpb_handle* h, double alpha, double beta, const double* d_daily, const double* d_weekly,
                   const double* d_monthly, int64_t n, double* d_out) {
    GPB_ENTER(h);
    if (n < 0) return gpb::set_error(h, -2, "post_blend: negative size");
    if (n > 0 && (!d_daily || !d_weekly || !d_monthly || !d_out)) return gpb::set_error(h, -2, "post_blend: null pointer");
    return gpb::post_blend(h, alpha, beta, d_daily, d_weekly, d_monthly, n, d_out);
}

int gpb_prep_returns(gpb_handle* h, const double* d_close, const double* d_open, int64_t T, int64_t A, int kind,
                     double* d_out) {
    GPB_ENTER(h);
    if (T < 0 || A < 0) return gpb::set_error(h, -2, "prep_returns: negative size");
    if (kind < 0 || kind > 2) return gpb::set_error(h, -2, "prep_returns: kind %d not in {0,1,2}", kind);
    if (T * A > 0 && (!d_close || !d_out)) return gpb::set_error(h, -2, "prep_returns: null pointer");
    if (kind == 1 && T * A > 0 && !d_open) return gpb::set_error(h, -2, "prep_returns: kind 1 needs the open series");
    return gpb::prep_returns(h, d_close, d_open, T, A, kind, d_out);
}

int gpb_prep_zscore(gpb_handle* h, const double* d_x, int64_t T, int64_t A, int ddof, double* d_out, int64_t ldo,
                    double* d_mean, double* d_std) {
    GPB_ENTER(h);
    if (T < 0 || A < 0) return gpb::set_error(h, -2, "prep_zscore: negative size");
    if (ddof < 0 || ddof > 1) return gpb::set_error(h, -2, "prep_zscore: ddof must be 0 or 1");
    if (T * A > 0 && !d_x) return gpb::set_error(h, -2, "prep_zscore: null pointer");
    if (d_out && ldo < A) return gpb::set_error(h, -2, "prep_zscore: ldo < A");
    return gpb::prep_zscore(h, d_x, T, A, ddof, d_out, ldo, d_mean, d_std);
}

int gpb_prep_windows(gpb_handle* h, const double* d_feat, const double* d_y, int64_t S, int64_t T, int D, int64_t N,
                     int64_t stride, double* d_X, double* d_Y) {
    GPB_ENTER(h);
    if (S < 0 || T < 0 || D < 1 || N < 1 || stride < 1) return gpb::set_error(h, -2, "prep_windows: bad sizes");
    if (S == 0 || T < N) return 0;  // no complete window
    if (!d_feat || !d_X) return gpb::set_error(h, -2, "prep_windows: null pointer");
    if ((d_y == nullptr) != (d_Y == nullptr)) return gpb::set_error(h, -2, "prep_windows: y and Y go together");
    return gpb::prep_windows(h, d_feat, d_y, S, T, D, N, stride, d_X, d_Y);
}

}  // extern "C"
