// Phase timing of the in-shared-memory 128x128 Cholesky + inverse (block_chol.cuh) with clock64().
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I. -o tools/leaf_prof tools/leaf_prof.cu
#include <cstdio>
#include <vector>
#include <cmath>
// per-warp phase accumulators of block_potrf_lower: [warp][0..4] = loop head, panel tile(s), barrier /
// tile-0 update, trailing update / diagonal factor, end-of-step sync wait
__shared__ long long prof_acc[16][5];
__shared__ int prof_step[16][16][5];   // the same per step (step index advances at stamp 4)
#define GPB_POTRF_DECL long long t_last_ = clock64(); int step_ = 0;
#define GPB_POTRF_STAMP(i) { if ((threadIdx.x & 31) == 0) { const long long t_now_ = clock64(); prof_acc[threadIdx.x >> 5][i] += t_now_ - t_last_; prof_step[threadIdx.x >> 5][step_ & 15][i] += (int)(t_now_ - t_last_); t_last_ = t_now_; } if (i == 4) ++step_; }
#include "../portfoliooptgp_b200/csrc/block_chol.cuh"
using namespace gpb;
namespace gpb {
__device__ __forceinline__ void potrf_timed(double* S, int np, int* fail, double* dinv, long long* tacc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int g = lane >> 2, q = lane & 3;
    if (tid == 0) *fail = 0;
    __syncthreads();
    for (int p = 0; p < np; p += 8) {
        double* M = dinv + (p >> 3) * 8 * DLD;
        long long ta = clock64();
        if (warp == 0) {
            // (a) all 32 lanes hold the whole lower triangle (broadcast loads) and run the same code
            double a[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) a[r][c] = S[(p + r) * SLD + p + c];
            double rs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double d = a[j][j];
                if (!(d > 0.0)) {
                    if (lane == 0 && *fail == 0) *fail = p + j + 1;
                    d = 1.0;
                }
                rs[j] = rsqrt(d);
                a[j][j] = d * rs[j];
#pragma unroll
                for (int r = j + 1; r < 8; ++r) a[r][j] *= rs[j];
#pragma unroll
                for (int r = j + 1; r < 8; ++r)
#pragma unroll
                    for (int k = j + 1; k <= r; ++k) a[r][k] = fma(-a[r][j], a[k][j], a[r][k]);
            }
            // inverse of the factor: lane c (mod 8) solves column c; 1/L_jj = rs[j]
            double x[8];
            const int c = lane & 7;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < r; ++k) v = fma(-a[r][k], x[k], v);
                x[r] = (r >= c) ? v * rs[r] : 0.0;
            }
            // write back: lane r (< 8) writes row r of L (static register indices via the unrolled select)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (lane == r) {
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) S[(p + r) * SLD + p + cc] = (cc <= r) ? a[r][cc] : 0.0;
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r) M[r * DLD + c] = x[r];
            }
        }
        __syncthreads();
        long long tb = clock64();
        const int m = np - p - 8;  // rows below the block
        if (m > 0) {
            const int mt = m >> 3;
            // (b) panel tile <- tile * M^T  (in place: a warp's operand loads complete before its stores)
            for (int ti = warp; ti < mt; ti += nwarps) {
                double* Pt = S + (p + 8 + ti * 8) * SLD + p;
                double c0 = 0.0, c1 = 0.0;
                warp_tile_mma(c0, c1, Pt, SLD, 1, M, 1, DLD, 8, 1.0);
                __syncwarp();
                *reinterpret_cast<double2*>(Pt + g * SLD + 2 * q) = make_double2(c0, c1);
            }
            __syncthreads();
            long long tc = clock64();
            // (c) trailing update, 8x8 tiles (ti >= tj) of the trailing matrix: C -= P_ti P_tj^T
            int idx = 0;
            for (int ti = 0; ti < mt; ++ti) {
                for (int tj = 0; tj <= ti; ++tj, ++idx) {
                    if (idx % nwarps != warp) continue;
                    double* C = S + (p + 8 + ti * 8 + g) * SLD + p + 8 + tj * 8 + 2 * q;
                    double2 cc = *reinterpret_cast<double2*>(C);
                    const double* Pi = S + (p + 8 + ti * 8) * SLD + p;
                    const double* Pj = S + (p + 8 + tj * 8) * SLD + p;
                    warp_tile_mma(cc.x, cc.y, Pi, SLD, 1, Pj, 1, SLD, 8, -1.0);
                    *reinterpret_cast<double2*>(C) = cc;
                }
            }
            __syncthreads();
            long long td = clock64();
            if (threadIdx.x == 0) { tacc[0] += tb - ta; tacc[1] += tc - tb; tacc[2] += td - tc; }
        }
    }
}

}

__global__ void __launch_bounds__(512) prof_kernel(const double* A, int n, long long* stamps, double* out) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm; double* T = sm + 128 * SLD; double* dinv = T + 64 * TLD; int* fail = (int*)(dinv + DINV_DOUBLES);
    const int tid = threadIdx.x, nt = blockDim.x;
    long long t0 = clock64();
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) S[i * SLD + j] = (j <= i) ? A[i * n + j] : 0.0; }
    __syncthreads();
    long long t1 = clock64();
    long long tacc[3] = {0, 0, 0};
    potrf_timed(S, n, fail, dinv, tacc);
    long long t2 = clock64();
    block_trtri_lower_inplace(S, n, T, dinv);
    long long t3 = clock64();
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) out[i * n + j] = S[i * SLD + j]; }
    __syncthreads();
    long long t4 = clock64();
    if (tid == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t3 - t2; stamps[3] = t4 - t3; stamps[4] = tacc[0]; stamps[5] = tacc[1]; stamps[6] = tacc[2]; }
}

// current production routines + isolated sub-phases (16 sequential diagonal factors; one full panel
// product; one full trailing update at p = 0)
template <int FUSED>
__global__ void __launch_bounds__(512) prof2_kernel(const double* A, int n, long long* stamps) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm; double* T = sm + 128 * SLD; double* dinv = T + 64 * TLD; int* fail = (int*)(dinv + DINV_DOUBLES);
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    const int g = lane >> 2, q = lane & 3;
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) S[i * SLD + j] = (j <= i) ? A[i * n + j] : 0.0; }
    if (tid < 80) (&prof_acc[0][0])[tid] = 0;
    for (int e = tid; e < 16 * 16 * 5; e += nt) (&prof_step[0][0][0])[e] = 0;
    __syncthreads();
    long long t0 = clock64();
    // 512 threads: the production pair (factor with the inverse built beside it, then W to the lower triangle);
    // otherwise the separate factor and recursive-doubling inverse
    if (FUSED) block_potrf_inv(S, n, fail, dinv, T); else block_potrf_lower(S, n, fail, dinv);
    long long t1 = clock64();
    if (tid < 80) stamps[8 + tid] = prof_acc[tid / 5][tid % 5];          // all warps
    for (int e = tid; e < 16 * 16 * 5; e += nt) stamps[128 + e] = (&prof_step[0][0][0])[e];
    if (FUSED) block_w_to_lower(S, n, dinv); else block_trtri_lower_inplace(S, n, T, dinv);
    __syncthreads();
    long long t2 = clock64();
    // reload, then 16 diagonal factors back to back on warp 0 (values are garbage after the first; timing only)
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) S[i * SLD + j] = (j <= i) ? A[i * n + j] : 0.0; }
    __syncthreads();
    long long t3 = clock64();
    if (warp == 0) for (int p = 0; p < n; p += 8) { warp_diag_factor(S, p, fail, dinv + (p >> 3) * 8 * DLD); __syncwarp(); }
    __syncthreads();
    long long t4 = clock64();
    {   // one panel product at p = 0
        const int p = 0; const double* M = dinv; const int mt = (n - 8) >> 3;
        for (int ti = warp; ti < mt; ti += nwarps) {
            double* Pt = S + (p + 8 + ti * 8) * SLD + p;
            double c0 = 0.0, c1 = 0.0;
            warp_tile_mma(c0, c1, Pt, SLD, 1, M, 1, DLD, 8, 1.0);
            __syncwarp();
            *reinterpret_cast<double2*>(Pt + g * SLD + 2 * q) = make_double2(c0, c1);
        }
    }
    __syncthreads();
    long long t5 = clock64();
    {   // one trailing update at p = 0, all warps
        const int mt = (n - 8) >> 3, ntiles = mt * (mt + 1) / 2;
        for (int t = warp; t < ntiles; t += nwarps) warp_trailing_tile(S, 0, t);
    }
    __syncthreads();
    long long t6 = clock64();
    for (int i = 0; i < 16; ++i) __syncthreads();
    long long t7 = clock64();
    if (tid == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t4 - t3; stamps[3] = t5 - t4; stamps[4] = t6 - t5; stamps[5] = (t7 - t6) / 16; }
}

// isolated micro-latencies
__global__ void lat_kernel(long long* out, double x) {
    double v = x; long long t0 = clock64();
    for (int i = 0; i < 64; ++i) v = rsqrt(v + 1.0);
    long long t1 = clock64();
    for (int i = 0; i < 64; ++i) v = 1.0 / (v + 1.5);
    long long t2 = clock64();
    double c0 = v, c1 = v;
    for (int i = 0; i < 64; ++i) dmma_8x8x4(c0, c1, v, x);
    long long t3 = clock64();
    for (int i = 0; i < 64; ++i) v = sqrt(v + 2.0);
    long long t4 = clock64();
    for (int i = 0; i < 64; ++i) v = __shfl_sync(0xffffffffu, v, (i * 7) & 31) + 1.0;
    long long t5 = clock64();
    for (int i = 0; i < 64; ++i) v = fma(v, 1.0000001, 0.5);
    long long t6 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 64; out[3] = (t4 - t3) / 64; out[4] = (t5 - t4) / 64; out[5] = (t6 - t5) / 64; out[6] = (long long)(v + c0 + c1); }
}

int main() {
    const int n = 128;
    std::vector<double> G(n * n), A(n * n);
    for (int i = 0; i < n * n; ++i) G[i] = sin(0.37 * i) ;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += G[i * n + k] * G[j * n + k]; A[i * n + j] = s / n + (i == j ? 0.5 : 0.0); }
    double *dA, *dO; long long* dS;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dO, n * n * 8); cudaMalloc(&dS, (128 + 1280) * 8);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    size_t smem = (128 * SLD + 64 * TLD + DINV_DOUBLES + 16) * 8;
    cudaFuncSetAttribute(prof_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int threads : {256, 512}) {
        for (int rep = 0; rep < 2; ++rep) prof_kernel<<<1, threads, smem>>>(dA, n, dS, dO);
        long long st[7]; cudaMemcpy(st, dS, 56, cudaMemcpyDeviceToHost);
        printf("threads %d cycles: load %lld potrf %lld [a %lld b %lld c %lld] trtri %lld store %lld (err %s)\n", threads, st[0], st[1], st[4], st[5], st[6], st[2], st[3], cudaGetErrorString(cudaGetLastError()));
    }
    cudaFuncSetAttribute(prof2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(prof2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int threads = cfg == 0 ? 256 : 512, fused = cfg == 2;
        for (int rep = 0; rep < 2; ++rep) {
            if (fused) prof2_kernel<1><<<1, threads, smem>>>(dA, n, dS); else prof2_kernel<0><<<1, threads, smem>>>(dA, n, dS);
        }
        long long st[6]; cudaMemcpy(st, dS, 48, cudaMemcpyDeviceToHost);
        printf("%s, threads %d cycles: potrf %lld %s %lld | 16 diag factors %lld, panel product(p=0) %lld, trailing(p=0) %lld, syncthreads %lld (err %s)\n",
               fused ? "fused factor + inverse" : "separate factor, inverse", threads, st[0], fused ? "W-to-lower" : "trtri", st[1], st[2], st[3], st[4], st[5],
               cudaGetErrorString(cudaGetLastError()));
        long long pa[80]; cudaMemcpy(pa, dS + 8, 640, cudaMemcpyDeviceToHost);
        printf("  potrf phases, warp 0: head %lld panel0 %lld tile0-update %lld diag %lld sync-wait %lld\n", pa[0], pa[1], pa[2], pa[3], pa[4]);
        for (int w = 1; w < threads / 32; ++w)
            printf("    warp %2d: head %lld panel %lld barrier %lld trailing %lld sync-wait %lld\n", w, pa[5 * w], pa[5 * w + 1], pa[5 * w + 2], pa[5 * w + 3], pa[5 * w + 4]);
        if (threads == 512) {
            std::vector<long long> ps(1280); cudaMemcpy(ps.data(), dS + 128, 1280 * 8, cudaMemcpyDeviceToHost);
            printf("    per step, warp 0 waiting (head + sync-wait) | chain (panel0 + tile0 + diag) | longest 'trailing' phase of the other warps:\n");
            for (int k = 0; k < 15; ++k) {
                long long mx = 0; int mw = 0;
                for (int w = 1; w < 16; ++w) { const long long v = ps[(w * 16 + k) * 5 + 3]; if (v > mx) { mx = v; mw = w; } }
                printf("      k=%2d  wait %5lld  chain %5lld  | max trailing %5lld (warp %d), panel phase of warp 1 %lld\n", k, ps[k * 5 + 0] + ps[k * 5 + 4],
                       ps[k * 5 + 1] + ps[k * 5 + 2] + ps[k * 5 + 3], mx, mw, ps[(16 + k) * 5 + 1]);
            }
        }
    }
    lat_kernel<<<1, 32>>>(dS, 1.3);
    long long l[7]; cudaMemcpy(l, dS, 56, cudaMemcpyDeviceToHost);
    printf("latency cycles: rsqrt+add %lld  div+add %lld  dmma-chain %lld  sqrt+add %lld  shfl+add %lld  dfma %lld\n", l[0], l[1], l[2], l[3], l[4], l[5]);
    return 0;
}
