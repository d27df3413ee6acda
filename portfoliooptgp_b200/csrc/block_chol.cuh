// block_chol.cuh -- CTA-cooperative dense kernels on a matrix resident in shared memory:
// Cholesky factorisation, triangular inverse, triangular solves.  Shared by the 128x128 leaf of
// the blocked Cholesky (cholesky.cu) and by the one-GP-per-CTA batched path (batched.cu)
// (north_star subsystems 2 and 4).  All routines are called by every thread of the CTA.
#pragma once
#include <cuda_runtime.h>

namespace gpb {

// In-place lower Cholesky of the n x n matrix S (row-major, stride ld, only the lower triangle is
// read).  Square-root-free elimination with ONE barrier per column: column j is kept unscaled
// while it eliminates, all columns are scaled by 1/sqrt(d_j) at the end.  On exit S holds L in
// its lower triangle; the strict upper triangle is zeroed.  *fail (shared) receives the 1-based
// index of the first non-positive pivot (0 = none); the factorisation continues with the pivot
// replaced by 1 so that no NaNs are produced.
__device__ __forceinline__ void block_potrf_lower(double* S, int ld, int n, int* fail) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tx = tid & 31, ty = tid >> 5, nwarps = nt >> 5;
    if (tid == 0) *fail = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        double d = S[j * ld + j];
        if (!(d > 0.0)) {
            // every thread sees the same value; thread 0 records and repairs
            if (tid == 0 && *fail == 0) *fail = j + 1;
            d = 1.0;
        }
        const double invd = 1.0 / d;
        // trailing update with the unscaled column: S[i][k] -= S[i][j] * S[k][j] / d, j < k <= i
        for (int i = j + 1 + ty; i < n; i += nwarps) {
            const double lij = S[i * ld + j] * invd;
            for (int k = j + 1 + tx; k <= i; k += 32) S[i * ld + k] = fma(-lij, S[k * ld + j], S[i * ld + k]);
        }
        __syncthreads();
    }
    // scale: L[i][j] = S[i][j] / sqrt(d_j) (i > j), L[j][j] = sqrt(d_j); zero the strict upper part
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, j = idx - i * n;
        if (j > i) {
            S[i * ld + j] = 0.0;
        }
    }
    __syncthreads();
    // diagonal last (columns are scaled by values derived from the diagonal)
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, j = idx - i * n;
        if (j < i) {
            double d = S[j * ld + j];
            if (!(d > 0.0)) d = 1.0;
            S[i * ld + j] *= rsqrt(d);
        }
    }
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        double d = S[j * ld + j];
        if (!(d > 0.0)) d = 1.0;
        S[j * ld + j] = sqrt(d);
    }
    __syncthreads();
}

__device__ __forceinline__ int packed_row(int i) { return i * (i + 1) / 2; }

// Wp (packed lower, row i at offset i(i+1)/2) = inverse of the lower-triangular L (row-major, ld).
// Row-oriented forward substitution: row i of W from rows < i; one barrier per row, each column's
// dot product split over 4 adjacent lanes.
__device__ __forceinline__ void block_trtri_lower_packed(const double* L, int ld, int n, double* Wp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int part = tid & 3, jbase = tid >> 2, jstride = nt >> 2;
    for (int i = 0; i < n; ++i) {
        const double inv = 1.0 / L[i * ld + i];
        for (int j0 = 0; j0 <= i; j0 += jstride) {
            const int j = j0 + jbase;
            double s = 0.0;
            if (j < i) {
                for (int k = j + part; k < i; k += 4) s = fma(L[i * ld + k], Wp[packed_row(k) + j], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0 && j <= i) Wp[packed_row(i) + j] = ((j == i) ? 1.0 : -s) * inv;
        }
        __syncthreads();
    }
}

}  // namespace gpb
