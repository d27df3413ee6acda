// batched.cu -- many small independent exact GPs, ONE GP PER CTA with its covariance matrix resident
// in shared memory (north_star subsystem 4; SURVEY.md 2.1 row K8, hard part H6).
//
// Replaces the sequential re-fit loops of the reference -- assets x rolling windows x restarts,
// Multi-Input_GPR/main.py:414-456 and models/model_trainer.py:26-48, each iteration a fresh
// gpflow.models.GPR + Scipy().minimize -- by one launch that evaluates B objectives (and gradients)
// at once.  Per GP: fused assembly -> blocked Cholesky (DMMA tiles) -> in-place inverse ->
// alpha, quadratic form, log-det -> K^-1 tiles formed on DMMA and consumed immediately by the
// fused gradient contraction (K^-1 and dK/dtheta are never stored).  HBM traffic per GP is
// 8(N D + N) bytes in and 8(2 + P) out; everything else stays on chip.
//
// N <= 128 (one 128 x 132 fp64 tile = 135 KB of the 227 KB shared memory), D <= 16.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "block_chol.cuh"
#include "engine.cuh"
#include "shapes.cuh"

namespace gpb {

constexpr int BT = 512;  // threads per CTA (16 warps: four per scheduler to hide the FP64 / DMMA / shared-memory latencies)
constexpr int BW = BT / 32;

struct BatchedSmem {
    // offsets in doubles
    static constexpr int S = 0;
    static constexpr int T = S + 128 * SLD;
    static constexpr int Y = T + 64 * TLD;       // y, a, alpha: 3 x 128
    static constexpr int RED = Y + 3 * 128;      // BW x (GPB_MAX_PARAMS + 2)
    static constexpr int MISC = RED + BW * (GPB_MAX_PARAMS + 2);  // scalars
    static constexpr int DINV = MISC + 32;       // inverted 8x8 diagonal blocks
    static constexpr int XS = DINV + DINV_DOUBLES;  // X tile: 128 x (DP + 1)  (odd stride: 2-way instead of 16-way bank conflicts)
};

template <int DP>
constexpr size_t batched_smem_bytes() { return (size_t)(BatchedSmem::XS + 128 * (DP + 1) + 16) * sizeof(double); }

// mode 0: LML only; 1: LML + gradient; 2: predict_f at Ns points per GP
template <int DP, bool FAST, class SH = DynShape>
__global__ void __launch_bounds__(BT, 1)
batched_gp_kernel(const DevKernel* __restrict__ kps, const int* __restrict__ kbad, const double* __restrict__ X,
                  const double* __restrict__ Yc, const double* __restrict__ noise, const int* __restrict__ nrows,
                  int Nmax, int D, int mode, double* __restrict__ out, int* __restrict__ info,
                  const double* __restrict__ Xs_new, int Ns, double* __restrict__ mean_out,
                  double* __restrict__ var_out, long long* __restrict__ prof) {
    using SHE = UnitWeights<SH>;   // element routines: the per-GP descriptor is in global memory
    extern __shared__ __align__(16) double sm[];
    double* S = sm + BatchedSmem::S;
    double* T = sm + BatchedSmem::T;
    double* ys = sm + BatchedSmem::Y;
    double* as = ys + 128;
    double* als = ys + 256;
    double* red = sm + BatchedSmem::RED;
    double* misc = sm + BatchedSmem::MISC;
    double* dinv = sm + BatchedSmem::DINV;
    int* fail = reinterpret_cast<int*>(misc + 8);
    double* xs = sm + BatchedSmem::XS;  // [128][XSTR]
    constexpr int XSTR = DP + 1;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    // ragged batches (expanding windows, Multi-Input_GPR/main.py:414-423): GP b uses the first nrows[b] of its
    // Nmax rows; the rest of its tile is identity padding like the round-up to a multiple of 8
    const int N = nrows ? min(max(nrows[b], 1), Nmax) : Nmax;
    const int np = (N + 7) & ~7;
    // per-GP kernel descriptor, built by build_dev_kernels_kernel; read-only global memory (not shared
    // memory) so that the compiler may keep its loop-invariant fields in registers
    const DevKernel& kp = kps[b];
    const int P = kp.n_params;
    const double* Xb = X + (size_t)b * Nmax * D;

    const bool do_prof = (prof != nullptr) && blockIdx.x == 0 && tid == 0;
    if (do_prof) prof[0] = clock64();
    if (tid == 0) misc[9] = (double)kbad[b];
    for (int e = tid; e < 128 * DP; e += BT) {
        const int r = e / DP, d = e % DP;
        xs[r * XSTR + d] = (r < N && d < D) ? Xb[r * D + d] : 0.0;
    }
    if (tid < 128) ys[tid] = (tid < N) ? Yc[(size_t)b * Nmax + tid] : 0.0;
    __syncthreads();
    const double nv = noise[b];
    if (do_prof) prof[1] = clock64();

    // ---- assembly: 2 x 2 blocks of the lower triangle, four elements advance together (kernel_value_2x2)
    const int nt8 = np >> 3;
    {
        const int nb2 = np >> 1;
        const bool fastk = SH::is_static || (kp.n_leaves <= GRAD_FAST_LEAVES);
        for (int t = tid; t < nb2 * (nb2 + 1) / 2; t += BT) {
            int bi, bj;
            tri_tile(t, bi, bj);
            const int i0 = 2 * bi, j0 = 2 * bj;
            double v[4];
            if (!SH::is_static && DP > 8) {
                // interpreter at D > 8: four coordinate vectors of 16 doubles do not fit the 128-register budget
                // of a 512-thread CTA (6 KB of spill code, lib/ptxas.log of round 1) -- one element at a time, its
                // two vectors re-read from shared memory
#pragma unroll 1
                for (int e = 0; e < 4; ++e) {
                    double xi[DP], xj[DP];
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        xi[d] = xs[(i0 + (e >> 1)) * XSTR + d];
                        xj[d] = xs[(j0 + (e & 1)) * XSTR + d];
                    }
                    v[e] = kernel_value<DP>(kp, xi, xj);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i0 + (e >> 1), j = j0 + (e & 1);
                    if (i >= N || j >= N) v[e] = (i == j) ? 1.0 : 0.0;   // identity padding
                    else if (i == j) v[e] += nv;
                }
                *reinterpret_cast<double2*>(S + i0 * SLD + j0) = make_double2(v[0], v[1]);
                *reinterpret_cast<double2*>(S + (i0 + 1) * SLD + j0) = make_double2(v[2], v[3]);
                continue;
            }
            double xa[DP], xb[DP], xj0[DP], xj1[DP];
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                xa[d] = xs[i0 * XSTR + d];
                xb[d] = xs[(i0 + 1) * XSTR + d];
                xj0[d] = xs[j0 * XSTR + d];
                xj1[d] = xs[(j0 + 1) * XSTR + d];
            }
            if (fastk) {
                kernel_value_2x2<DP, SHE>(kp, xa, xb, xj0, xj1, v);
            } else {
                v[0] = kernel_value<DP>(kp, xa, xj0);
                v[1] = kernel_value<DP>(kp, xa, xj1);
                v[2] = kernel_value<DP>(kp, xb, xj0);
                v[3] = kernel_value<DP>(kp, xb, xj1);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + (e >> 1), j = j0 + (e & 1);
                if (i >= N || j >= N) v[e] = (i == j) ? 1.0 : 0.0;   // identity padding
                else if (i == j) v[e] += nv;
            }
            *reinterpret_cast<double2*>(S + i0 * SLD + j0) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(S + (i0 + 1) * SLD + j0) = make_double2(v[2], v[3]);
        }
    }
    __syncthreads();
    if (do_prof) prof[2] = clock64();

    block_potrf_inv(S, np, fail, dinv, T);   // L below the diagonal, W = L^-1 built beside it (block_chol.cuh)
    if (do_prof) prof[3] = clock64();
    // log-det (fixed order) by warp 0 from the inverted diagonal blocks, beside the move of W by the other warps
    if (warp == 0) {
        const double s = warp_logdiag_from_dinv(dinv, N);
        if (lane == 0) misc[10] = s;
    }
    block_w_to_lower(S, np, dinv, 1);            // S <- W = L^-1, row-major lower triangle
    if (do_prof) prof[4] = clock64();

    // ---- a = W y, alpha = W^T a: matrix-vector products on DMMA tiles, the vector broadcast over the eight
    // columns of the B operand (every column of the result tile holds the answer; column 0 is kept).  A warp per
    // 8-row tile of W resp. 8-column tile of W^T, two accumulator chains: ~450 cycles each, where a warp per row
    // with a shuffle reduction took 4 K and a thread per column 1.3 K.
    for (int ti = warp; ti < nt8; ti += BW) {
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        const double* ap = S + (ti * 8 + g) * SLD + q;
        for (int kk = 0; kk < (ti + 1) * 8; kk += 8) {
            dmma_8x8x4(c0, c1, ap[kk], ys[kk + q]);
            dmma_8x8x4(d0, d1, ap[kk + 4], ys[kk + 4 + q]);
        }
        if (q == 0) as[ti * 8 + g] = c0 + d0;
    }
    __syncthreads();
    for (int tj = warp; tj < nt8; tj += BW) {
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        const double* ap = S + q * SLD + tj * 8 + g;          // A(m, k) = W[k][tj 8 + m]
        for (int kk = tj * 8; kk < np; kk += 8) {
            dmma_8x8x4(c0, c1, ap[kk * SLD], as[kk + q]);
            dmma_8x8x4(d0, d1, ap[(kk + 4) * SLD], as[kk + 4 + q]);
        }
        if (q == 0) als[tj * 8 + g] = c0 + d0;
    }
    if (warp == 1) {
        double s = 0.0;
        for (int i = lane; i < N; i += 32) s = fma(as[i], as[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) misc[11] = s;
    }
    __syncthreads();

    if (do_prof) prof[5] = clock64();
    if (mode == 2) {
        // predict_f: mean_s = k_s^T alpha ; var_s = k_ss - |W k_s|^2.  One warp per test point.
        double* ks = T;  // BW x 128 scratch
        for (int s0 = warp; s0 < Ns; s0 += BW) {
            double xn[DP];
#pragma unroll
            for (int d = 0; d < DP; ++d) xn[d] = (d < D) ? Xs_new[((size_t)b * Ns + s0) * D + d] : 0.0;
            double m = 0.0;
            for (int i = lane; i < np; i += 32) {
                double xi[DP];
#pragma unroll
                for (int d = 0; d < DP; ++d) xi[d] = xs[i * XSTR + d];
                const double kv = (i < N) ? kernel_value_auto<DP, SHE>(kp, xi, xn) : 0.0;
                ks[warp * 128 + i] = kv;
                m = fma(kv, als[i], m);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
            __syncwarp();
            double ss = 0.0;
            for (int i = lane; i < np; i += 32) {
                double v = 0.0;
                for (int j = 0; j <= i; ++j) v = fma(S[i * SLD + j], ks[warp * 128 + j], v);
                ss = fma(v, v, ss);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) {
                mean_out[(size_t)b * Ns + s0] = m;
                var_out[(size_t)b * Ns + s0] = (SH::is_static ? kernel_value_fast<DP, SHE>(kp, xn, xn) : kernel_value<DP>(kp, xn, xn)) - ss;
            }
            __syncwarp();
        }
        if (tid == 0) info[b] = (misc[9] != 0.0) ? -(int)misc[9] : *fail;
        return;
    }

    double* o = out + (size_t)b * (2 + P);
    if (mode == 1) {
        // ---- gradient: K^-1 tile = sum_{k >= ti*8} W[k, ti-blk]^T W[k, tj-blk] on DMMA, consumed in place
        constexpr bool fast = FAST;   // register accumulators (<= 4 leaves, no ARD) vs generic path: two kernels
        GradAcc A;
        A.zero();
        double tr = 0.0;
        double acc[FAST ? 1 : GPB_MAX_PARAMS + 1];
        if (!fast)
            for (int p = 0; p <= P; ++p) acc[p] = 0.0;
        for (int t = warp; t < nt8 * (nt8 + 1) / 2; t += BW) {
            {
                int ti, tj;
                tri_tile(t, ti, tj);
                double c0 = 0.0, c1 = 0.0;
                const double* Wk = S + (ti * 8) * SLD;
                warp_tile_mma(c0, c1, Wk + ti * 8, 1, SLD, Wk + tj * 8, SLD, 1, np - ti * 8, 1.0);
                const int i = ti * 8 + g, j0 = tj * 8 + 2 * q;
                if (i < N) {
                    double xi[DP], xj[DP];
#pragma unroll
                    for (int d = 0; d < DP; ++d) xi[d] = xs[i * XSTR + d];
                    const double ai = als[i];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int j = j0 + c;
                        if (j <= i) {
                            double w = ai * als[j] - (c == 0 ? c0 : c1);
                            if (j == i) tr += w; else w *= 2.0;
#pragma unroll
                            for (int d = 0; d < DP; ++d) xj[d] = xs[j * XSTR + d];
                            if (fast) kernel_value_grad_fast<DP, SHE>(kp, xi, xj, w, A);
                            else kernel_value_grad<DP>(kp, xi, xj, w, acc);
                        }
                    }
                }
                __syncwarp();
            }
        }
        double* redw = red + warp * (GPB_MAX_PARAMS + 2);
        if (fast) {
            for (int p = lane; p <= P; p += 32) redw[p] = 0.0;
            __syncwarp();
            grad_flush<SH>(kp, A, redw);
        } else {
            for (int p = 0; p < P; ++p) {
                double v = acc[p];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) redw[p] = v;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tr += __shfl_down_sync(0xffffffffu, tr, off);
        if (lane == 0) redw[P] = tr;
        __syncthreads();
        if (tid <= P) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < BW; ++w) v += red[w * (GPB_MAX_PARAMS + 2) + tid];
            if (tid == P) o[1] = 0.5 * v; else o[2 + tid] = 0.5 * v;
        }
    }
    if (do_prof) prof[6] = clock64();
    if (tid == 0) {
        o[0] = -0.5 * misc[11] - 0.5 * (double)N * 1.8378770664093453 - misc[10];   // log(2 pi)
        info[b] = (misc[9] != 0.0) ? -(int)misc[9] : *fail;
    }
}

// one thread per GP: spec + theta[b] -> DevKernel[b]
static __global__ void build_dev_kernels_kernel(const __grid_constant__ gpb_kernel_spec spec, const double* __restrict__ theta,
                                         int64_t B, DevKernel* __restrict__ out, int* __restrict__ bad) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bad[b] = build_dev_kernel_core(spec, theta + b * spec.n_params, out + b);
}

template <int DP, bool FAST, class SH = DynShape>
static int launch_batched_dp(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                             const double* d_noise, const int* d_nrows, int64_t B, int N, int D, int mode, double* d_out,
                             int* d_info, const double* d_Xs, int Ns, double* d_mean, double* d_var) {
    auto kern = batched_gp_kernel<DP, FAST, SH>;
    constexpr size_t SMEM = batched_smem_bytes<DP>();
    static bool attr_set[GPB_MAX_DEVICES] = {};
    {
        cudaError_t e = ensure_dyn_smem(attr_set, h->device, kern, SMEM);
        if (e != cudaSuccess) return check_cuda(h, e, "batched cudaFuncSetAttribute");
    }
    static long long* d_prof = nullptr;   // GPB_BATCHED_PROF=1: phase cycle stamps of CTA 0 to stderr (debug)
    static int want_prof = -1;
    if (want_prof < 0) {
        const char* e = getenv("GPB_BATCHED_PROF");
        want_prof = (e && e[0] == '1') ? 1 : 0;
        if (want_prof) cudaMalloc(&d_prof, 8 * sizeof(long long));
    }
    const size_t kbytes = ((size_t)B * sizeof(DevKernel) + 255) / 256 * 256;
    double* kbuf = workspace(h, BUF_AUX, kbytes + (size_t)B * sizeof(int));
    if (!kbuf) return -1;
    DevKernel* kps = reinterpret_cast<DevKernel*>(kbuf);
    int* kbad = reinterpret_cast<int*>(reinterpret_cast<char*>(kbuf) + kbytes);
    {
        ProfScope prof(h, PROF_BATCHED, h->stream);
        build_dev_kernels_kernel<<<(unsigned)((B + 127) / 128), 128, 0, h->stream>>>(h->spec, d_theta, B, kps, kbad);
        kern<<<(unsigned)B, BT, SMEM, h->stream>>>(kps, kbad, d_X, d_Yc, d_noise, d_nrows, N, D, mode, d_out, d_info, d_Xs, Ns,
                                                    d_mean, d_var, want_prof ? d_prof : nullptr);
        h->launches += 1;
    }
    if (want_prof && d_prof) {
        long long st[8];
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(st, d_prof, sizeof(st), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[batched prof, cycles] setup %lld assemble %lld potrf %lld trtri %lld vectors %lld grad/predict %lld total %lld\n",
                st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[6] - st[0]);
    }
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "batched_gp_kernel launch");
}


// launchers instantiated in other translation units (compile time: one object per kernel family)
#define GPB_BATCHED_PARAMS gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta, const double* d_noise, \
    const int* d_nrows, int64_t B, int N, int D, int mode, double* d_out, int* d_info, const double* d_Xs, int Ns, double* d_mean, \
    double* d_var
#define GPB_BATCHED_ARGS h, d_X, d_Yc, d_theta, d_noise, d_nrows, B, N, D, mode, d_out, d_info, d_Xs, Ns, d_mean, d_var
int launch_batched_generic(int dp, GPB_BATCHED_PARAMS);          // batched_generic.cu: interpreter, any expression (ARD, > 4 leaves)
int launch_batched_static(int shape, int dp, GPB_BATCHED_PARAMS); // batched_shapes.cu: straight-line shapes; -100 = no such instantiation

}  // namespace gpb
