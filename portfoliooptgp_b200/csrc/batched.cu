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
#include "block_chol.cuh"
#include "engine.cuh"

namespace gpb {

constexpr int BT = 256;  // threads per CTA
constexpr int BW = BT / 32;

struct BatchedSmem {
    // offsets in doubles
    static constexpr int S = 0;
    static constexpr int T = S + 128 * SLD;
    static constexpr int Y = T + 64 * TLD;       // y, a, alpha: 3 x 128
    static constexpr int RED = Y + 3 * 128;      // BW x (GPB_MAX_PARAMS + 2)
    static constexpr int MISC = RED + BW * (GPB_MAX_PARAMS + 2);  // scalars
    static constexpr int DINV = MISC + 32;       // inverted 8x8 diagonal blocks
    static constexpr int KP = DINV + DINV_DOUBLES;  // DevKernel
    static constexpr int XS = KP + (int)((sizeof(DevKernel) + 7) / 8);  // X tile: 128 x DP
};

template <int DP>
constexpr size_t batched_smem_bytes() { return (size_t)(BatchedSmem::XS + 128 * DP + 16) * sizeof(double); }

// mode 0: LML only; 1: LML + gradient; 2: predict_f at Ns points per GP
template <int DP>
__global__ void __launch_bounds__(BT, 1)
batched_gp_kernel(const __grid_constant__ gpb_kernel_spec spec, const double* __restrict__ X,
                  const double* __restrict__ Yc, const double* __restrict__ theta, const double* __restrict__ noise,
                  int N, int D, int mode, double* __restrict__ out, int* __restrict__ info,
                  const double* __restrict__ Xs_new, int Ns, double* __restrict__ mean_out,
                  double* __restrict__ var_out) {
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
    DevKernel& kp = *reinterpret_cast<DevKernel*>(sm + BatchedSmem::KP);
    double* xs = sm + BatchedSmem::XS;  // [128][DP]

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int np = (N + 7) & ~7;
    const int P = spec.n_params;
    const double* Xb = X + (size_t)b * N * D;
    const double* th = theta + (size_t)b * P;

    if (tid == 0) {
        int bad = build_dev_kernel_core(spec, th, &kp);
        misc[9] = (double)bad;
    }
    for (int e = tid; e < 128 * DP; e += BT) {
        const int r = e / DP, d = e % DP;
        xs[e] = (r < N && d < D) ? Xb[r * D + d] : 0.0;
    }
    if (tid < 128) ys[tid] = (tid < N) ? Yc[(size_t)b * N + tid] : 0.0;
    __syncthreads();
    const double nv = noise[b];

    // ---- assembly: 8x8 tiles of the lower triangle, fragment layout of the DMMA C tile
    const int nt8 = np >> 3;
    {
        for (int t = warp; t < nt8 * (nt8 + 1) / 2; t += BW) {
            {
                int ti, tj;
                tri_tile(t, ti, tj);
                const int i = ti * 8 + g, j0 = tj * 8 + 2 * q;
                double xi[DP], xj[DP];
#pragma unroll
                for (int d = 0; d < DP; ++d) xi[d] = xs[i * DP + d];
                double v[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int j = j0 + c;
#pragma unroll
                    for (int d = 0; d < DP; ++d) xj[d] = xs[j * DP + d];
                    double val;
                    if (i < N && j < N) {
                        val = kernel_value<DP>(kp, xi, xj);
                        if (i == j) val += nv;
                    } else {
                        val = (i == j) ? 1.0 : 0.0;  // identity padding
                    }
                    v[c] = val;
                }
                *reinterpret_cast<double2*>(S + i * SLD + j0) = make_double2(v[0], v[1]);
            }
        }
    }
    __syncthreads();

    block_potrf_lower(S, np, fail, dinv);
    // log-det (fixed order) by warp 0
    if (warp == 0) {
        double s = 0.0;
        for (int i = lane; i < N; i += 32) s += log(S[i * SLD + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) misc[10] = s;
    }
    __syncthreads();
    block_trtri_lower_inplace(S, np, T, dinv);   // S <- W = L^-1

    // ---- a = W y (warp per row), alpha = W^T a (thread per column)
    for (int i = warp; i < np; i += BW) {
        double s = 0.0;
        for (int j = lane; j <= i; j += 32) s = fma(S[i * SLD + j], ys[j], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) as[i] = s;
    }
    __syncthreads();
    if (tid < np) {
        double s = 0.0;
        for (int i = tid; i < np; ++i) s = fma(S[i * SLD + tid], as[i], s);
        als[tid] = s;
    }
    if (warp == 1) {
        double s = 0.0;
        for (int i = lane; i < N; i += 32) s = fma(as[i], as[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) misc[11] = s;
    }
    __syncthreads();

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
                for (int d = 0; d < DP; ++d) xi[d] = xs[i * DP + d];
                const double kv = (i < N) ? kernel_value<DP>(kp, xi, xn) : 0.0;
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
                var_out[(size_t)b * Ns + s0] = kernel_value<DP>(kp, xn, xn) - ss;
            }
            __syncwarp();
        }
        if (tid == 0) info[b] = (misc[9] != 0.0) ? -(int)misc[9] : *fail;
        return;
    }

    double* o = out + (size_t)b * (2 + P);
    if (mode == 1) {
        // ---- gradient: K^-1 tile = sum_{k >= ti*8} W[k, ti-blk]^T W[k, tj-blk] on DMMA, consumed in place
        const bool fast = grad_fast_ok(kp);
        GradAcc A;
        A.zero();
        double tr = 0.0;
        double acc[GPB_MAX_PARAMS + 1];
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
                    for (int d = 0; d < DP; ++d) xi[d] = xs[i * DP + d];
                    const double ai = als[i];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int j = j0 + c;
                        if (j <= i) {
                            double w = ai * als[j] - (c == 0 ? c0 : c1);
                            if (j == i) tr += w; else w *= 2.0;
#pragma unroll
                            for (int d = 0; d < DP; ++d) xj[d] = xs[j * DP + d];
                            if (fast) kernel_value_grad_fast<DP>(kp, xi, xj, w, A);
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
            grad_flush(kp, A, redw);
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
    if (tid == 0) {
        o[0] = -0.5 * misc[11] - 0.5 * (double)N * 1.8378770664093453 - misc[10];   // log(2 pi)
        info[b] = (misc[9] != 0.0) ? -(int)misc[9] : *fail;
    }
}

template <int DP>
static int launch_batched_dp(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                             const double* d_noise, int64_t B, int N, int D, int mode, double* d_out, int* d_info,
                             const double* d_Xs, int Ns, double* d_mean, double* d_var) {
    auto kern = batched_gp_kernel<DP>;
    constexpr size_t SMEM = batched_smem_bytes<DP>();
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return check_cuda(h, e, "batched cudaFuncSetAttribute");
        attr_set = true;
    }
    ProfScope prof(h, PROF_BATCHED, h->stream);
    kern<<<(unsigned)B, BT, SMEM, h->stream>>>(h->spec, d_X, d_Yc, d_theta, d_noise, N, D, mode, d_out, d_info, d_Xs, Ns,
                                                d_mean, d_var);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "batched_gp_kernel launch");
}

int launch_batched(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta, const double* d_noise,
                   int64_t B, int64_t N, int D, int mode, double* d_out, int* d_info, const double* d_Xs, int64_t Ns,
                   double* d_mean, double* d_var) {
    if (!h->has_spec) return set_error(h, -3, "batched: no kernel set (gpb_set_kernel)");
    if (B <= 0) return 0;
    if (N < 1 || N > 128) return set_error(h, -2, "batched: N=%lld outside [1,128] (one GP per CTA in shared memory)", (long long)N);
    if (D != h->spec.n_dims) return set_error(h, -2, "batched: kernel expects D=%d, got %d", h->spec.n_dims, D);
    if (B > 0x7fffffffLL) return set_error(h, -2, "batched: B too large");
    int dp = 1;
    while (dp < D) dp <<= 1;
    switch (dp) {
        case 1: return launch_batched_dp<1>(h, d_X, d_Yc, d_theta, d_noise, B, (int)N, D, mode, d_out, d_info, d_Xs, (int)Ns, d_mean, d_var);
        case 2: return launch_batched_dp<2>(h, d_X, d_Yc, d_theta, d_noise, B, (int)N, D, mode, d_out, d_info, d_Xs, (int)Ns, d_mean, d_var);
        case 4: return launch_batched_dp<4>(h, d_X, d_Yc, d_theta, d_noise, B, (int)N, D, mode, d_out, d_info, d_Xs, (int)Ns, d_mean, d_var);
        case 8: return launch_batched_dp<8>(h, d_X, d_Yc, d_theta, d_noise, B, (int)N, D, mode, d_out, d_info, d_Xs, (int)Ns, d_mean, d_var);
        default: return launch_batched_dp<16>(h, d_X, d_Yc, d_theta, d_noise, B, (int)N, D, mode, d_out, d_info, d_Xs, (int)Ns, d_mean, d_var);
    }
}

}  // namespace gpb
