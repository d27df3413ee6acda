// svgp.cu -- sparse variational GP (north_star subsystem 5): minibatch ELBO data term, its
// hand-derived gradient, the KL term, and predict_f.
//
// Replaces gpflow.models.SVGP.elbo / predict_f as driven at test_scripts/SVGP.py:515-540:
// gpflow/models/svgp.py (elbo), gpflow/conditionals/conditionals.py + util.py
// (base_conditional_with_lm, white=True, q_sqrt [1,M,M]), gpflow/covariances (Kuu + jitter I, Kuf),
// gpflow/likelihoods Gaussian.variational_expectations, gpflow/kullback_leiblers.py gauss_kl
// (SURVEY.md 8a G13-G14).  TF autodiff of that graph is replaced by the adjoint below.
//
// Forward (whitened, one latent):            Backward of S = sum_b ve_b:
//   Lm = chol(k(Z,Z) + jitter I), Wm = Lm^-1    r_b = (y_b - fmean_b)/s2,  gv = -1/(2 s2)
//   A  = Wm k(Z, Xb)               [M,B]        gq_mu = A r
//   C  = Lq^T A                    [M,B]        gLq   = 2 gv A C^T                     (lower)
//   fmean_b = A[:,b].q_mu                       Abar  = q_mu r^T + 2 gv (Lq C - A)
//   fvar_b  = kdiag_b - |A[:,b]|^2 + |C[:,b]|^2 Kuf_bar = Wm^T Abar
//   ve_b = -log(2 pi s2)/2                      Lm_bar  = -tril(Kuf_bar A^T)
//          - ((y_b - fmean_b)^2 + fvar_b)/(2 s2)  Kuu_bar = Wm^T sym(tril(Lm^T Lm_bar)) Wm / 2
//                                               theta, Z gradients: fused contractions of Kuf_bar,
//                                               Kuu_bar, gv with dk/dtheta, dk/dz (never stored)
// All six M x M x B products and the M^3 ones run in dgemm.cu (DMMA): ~6 M^2 B flop per step.
//
// Flat result record (one device buffer so that data-parallel ranks all-reduce it in one call):
//   [0] S = sum_b ve_b   [1] dS/dnoise   [2, 2+P) dS/dtheta   then dS/dZ [M,D], dS/dq_mu [M],
//   dS/dq_sqrt [M,M] (row-major, lower triangle meaningful).
#include <math.h>

#include "engine.cuh"
#include "shapes.cuh"

namespace gpb {

static inline int64_t rup(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- fused contraction of a weight matrix with dk/dtheta and dk/d(first argument) --------------------
constexpr int CG_TILE = 64;
constexpr int CG_THREADS = 256;

// theta_part[block][p] = sum over the block's tiles of Wt[i][j] dk(a_i, b_j)/dtheta_p
// zbar_part[split][i][d] = zscale * sum_j Wt[i][j] dk(a_i, b_j)/da_i[d]
template <int DP, class SH = DynShape>
__global__ void __launch_bounds__(CG_THREADS)
cross_grad_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ Arows, int Na,
                  const double* __restrict__ Bcols, int Nb, int D, const double* __restrict__ Wt, int64_t ldw,
                  double zscale, double* __restrict__ theta_part, double* __restrict__ zbar_part) {
    __shared__ double wst[CG_TILE * (CG_TILE + 1)];   // staged weight tile; reused for the row reduction
    __shared__ double xb[CG_TILE][DP];
    __shared__ double red[CG_THREADS / 32][GPB_MAX_PARAMS + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * CG_TILE;
    const int nsplit = gridDim.y, split = blockIdx.y;
    const int r = (warp & 1) * 32 + lane;          // row within the tile
    const int cq = (warp >> 1) * 16;               // column quarter
    const int gi = row0 + r;
    double xi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) xi[d] = (gi < Na && d < D) ? Arows[(int64_t)gi * D + d] : 0.0;
    GradAcc A;
    A.zero();
    double gx[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) gx[d] = 0.0;
    const int ntiles = (Nb + CG_TILE - 1) / CG_TILE;
    for (int ct = split; ct < ntiles; ct += nsplit) {
        const int col0 = ct * CG_TILE;
        __syncthreads();
        for (int e = tid; e < CG_TILE * CG_TILE; e += CG_THREADS) {
            const int rr = e >> 6, cc = e & 63;
            const int64_t gr = row0 + rr, gc = col0 + cc;
            wst[rr * (CG_TILE + 1) + cc] = (gr < Na && gc < Nb) ? Wt[gr * ldw + gc] : 0.0;
        }
        for (int e = tid; e < CG_TILE * DP; e += CG_THREADS) {
            const int rr = e / DP, d = e % DP;
            const int64_t gc = col0 + rr;
            xb[rr][d] = (gc < Nb && d < D) ? Bcols[gc * D + d] : 0.0;
        }
        __syncthreads();
        if (gi < Na) {
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {
                const int cc = cq + c;
                if (col0 + cc >= Nb) break;
                double xj[DP];
#pragma unroll
                for (int d = 0; d < DP; ++d) xj[d] = xb[cc][d];
                const double w = wst[r * (CG_TILE + 1) + cc];
                kernel_value_grad_x_fast<DP, UnitWeights<SH>>(kp, xi, xj, w, A, gx);
            }
        }
    }
    __syncthreads();
    // rows: 4 warps share a row (one per column quarter): reduce through shared memory, fixed order
    double* zred = wst;  // [4][64][DP] <= 64*65 doubles for DP <= 16
#pragma unroll
    for (int d = 0; d < DP; ++d) zred[((warp >> 1) * CG_TILE + r) * DP + d] = gx[d];
    const int P = kp.n_params;
    for (int p = lane; p < P; p += 32) red[warp][p] = 0.0;
    __syncwarp();
    grad_flush<SH>(kp, A, red[warp]);
    __syncthreads();
    for (int e = tid; e < CG_TILE * DP; e += CG_THREADS) {
        const int rr = e / DP, d = e % DP;
        if (row0 + rr < Na) {
            double s = 0.0;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) s += zred[(qd * CG_TILE + rr) * DP + d];
            zbar_part[((int64_t)split * Na + row0 + rr) * DP + d] = zscale * s;
        }
    }
    if (tid < P) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < CG_THREADS / 32; ++w) s += red[w][tid];
        theta_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * GPB_MAX_PARAMS + tid] = s;
    }
}

// out[p] (+)= scale * sum_b part[b][p]   (one block per p, fixed order)
__global__ void reduce_rows_kernel(const double* __restrict__ part, int64_t nrows, int stride, double scale,
                                   double* __restrict__ out, int accumulate) {
    __shared__ double sm[256];
    const int p = blockIdx.x;
    double v = 0.0;
    for (int64_t b = threadIdx.x; b < nrows; b += 256) v += part[b * stride + p];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p] = (accumulate ? out[p] : 0.0) + scale * sm[0];
}

// gZ[i][d] (+)= sum_s zbar_part[s][i][d]   (DP-strided partials -> D-strided output)
__global__ void reduce_zbar_kernel(const double* __restrict__ part, int nsplit, int Na, int DP, int D,
                                   double* __restrict__ out, int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Na * D) return;
    const int i = e / D, d = e % D;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += part[((int64_t)sp * Na + i) * DP + d];
    out[e] = (accumulate ? out[e] : 0.0) + s;
}

// theta gradient of sum_b gv * k(x_b, x_b)
template <int DP>
__global__ void kdiag_grad_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ X, int N, int D,
                                  double gv, double* __restrict__ theta_part) {
    __shared__ double red[8][GPB_MAX_PARAMS + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GradAcc A;
    A.zero();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double xi[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) xi[d] = (d < D) ? X[i * D + d] : 0.0;
        kernel_value_grad_fast<DP>(kp, xi, xi, gv, A);
    }
    const int P = kp.n_params;
    for (int p = lane; p < P; p += 32) red[warp][p] = 0.0;
    __syncwarp();
    grad_flush(kp, A, red[warp]);
    __syncthreads();
    if (tid < P) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w][tid];
        theta_part[(int64_t)blockIdx.x * GPB_MAX_PARAMS + tid] = s;
    }
}

// ---- column statistics of A and C ------------------------------------------------------------------------
constexpr int CS_CH = 128;
__global__ void colstats_partial_kernel(const double* __restrict__ A, const double* __restrict__ C, int64_t ld, int M,
                                        int B, const double* __restrict__ qmu, double* __restrict__ p_mean,
                                        double* __restrict__ p_sa, double* __restrict__ p_sc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (b >= B) return;
    const int m0 = c * CS_CH, m1 = min(M, m0 + CS_CH);
    double fm = 0.0, sa = 0.0, sc = 0.0;
    for (int m = m0; m < m1; ++m) {
        const double a = A[(int64_t)m * ld + b];
        fm = fma(a, qmu[m], fm);
        sa = fma(a, a, sa);
        if (C) {
            const double cc = C[(int64_t)m * ld + b];
            sc = fma(cc, cc, sc);
        }
    }
    p_mean[(int64_t)c * B + b] = fm;
    p_sa[(int64_t)c * B + b] = sa;
    p_sc[(int64_t)c * B + b] = sc;
}

// per column: fmean, fvar; with y: r_b, ve_b and the noise-derivative term; block partials of the sums
__global__ void colstats_finish_kernel(const double* __restrict__ p_mean, const double* __restrict__ p_sa,
                                       const double* __restrict__ p_sc, int nch, int B, const double* __restrict__ kdiag,
                                       const double* __restrict__ y, double s2, double* __restrict__ fmean,
                                       double* __restrict__ fvar, double* __restrict__ rvec,
                                       double* __restrict__ blk_sums /* [gridDim.x][2] */) {
    __shared__ double sm[2][256];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double ve = 0.0, dn = 0.0;
    if (b < B) {
        double fm = 0.0, sa = 0.0, sc = 0.0;
        for (int c = 0; c < nch; ++c) {
            fm += p_mean[(int64_t)c * B + b];
            sa += p_sa[(int64_t)c * B + b];
            sc += p_sc[(int64_t)c * B + b];
        }
        const double fv = kdiag[b] - sa + sc;
        fmean[b] = fm;
        fvar[b] = fv;
        if (y) {
            const double res = y[b] - fm;
            rvec[b] = res / s2;
            const double q = res * res + fv;
            ve = -0.5 * 1.8378770664093453 - 0.5 * log(s2) - 0.5 * q / s2;
            dn = -0.5 / s2 + 0.5 * q / (s2 * s2);
        }
    }
    sm[0][threadIdx.x] = ve;
    sm[1][threadIdx.x] = dn;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + s];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blk_sums) {
        blk_sums[2 * blockIdx.x] = sm[0][0];
        blk_sums[2 * blockIdx.x + 1] = sm[1][0];
    }
}

// Abar = q_mu r^T + 2 gv (G - A), in place on G
__global__ void abar_kernel(double* __restrict__ G, const double* __restrict__ A, int64_t ld, int M, int B,
                            const double* __restrict__ qmu, const double* __restrict__ rvec, double gv2) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (b >= B) return;
    const int64_t o = (int64_t)m * ld + b;
    G[o] = fma(qmu[m], rvec[b], gv2 * (G[o] - A[o]));
}

// out[m] = sum_b A[m][b] v[b]  (warp per row)
__global__ void rowdot_kernel(const double* __restrict__ A, int64_t ld, int M, int B, const double* __restrict__ v,
                              double* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const double* a = A + (int64_t)row * ld;
    double s = 0.0;
    for (int b = lane; b < B; b += 32) s = fma(a[b], v[b], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

// strict upper triangle <- 0 (in place), [n, ld]
__global__ void zero_strict_upper_kernel(double* __restrict__ A, int64_t ld, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j < n && j > i) A[(int64_t)i * ld + j] = 0.0;
}

// Sym = sym-from-lower(Q) (diagonal kept): the P + P^T of the Cholesky adjoint with P = tril(Q), diag halved
__global__ void sym_from_lower_kernel(const double* __restrict__ Q, int64_t ldq, double* __restrict__ S, int64_t lds, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    S[(int64_t)i * lds + j] = (j <= i) ? Q[(int64_t)i * ldq + j] : Q[(int64_t)j * ldq + i];
}

// copy the lower triangle of src into dst (row-major [n, ld]); zero above the diagonal
__global__ void copy_lower_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    dst[(int64_t)i * ldd + j] = (j <= i) ? src[(int64_t)i * lds + j] : 0.0;
}

// KL[q || N(0,I)] (whitened) and, in place, g <- scale * g - dKL (for the q_mu and q_sqrt slots).
// Single block, fixed-order reduction.  out2[0] = KL.
__global__ void kl_finish_kernel(const double* __restrict__ qmu, const double* __restrict__ Lq, int64_t ldq, int M,
                                 double scale, double* __restrict__ g_head, int n_head, double* __restrict__ g_qmu,
                                 double* __restrict__ g_lq, double* __restrict__ out2, int apply) {
    __shared__ double sm[1024];
    double acc = 0.0;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (apply)
        for (int i = tid; i < n_head; i += nt) g_head[i] *= scale;
    for (int i = tid; i < M; i += nt) {
        const double m = qmu[i];
        acc = fma(m, m, acc);
        if (apply) g_qmu[i] = scale * g_qmu[i] - m;
    }
    for (int64_t e = tid; e < (int64_t)M * M; e += nt) {
        const int i = (int)(e / M), j = (int)(e % M);
        if (j <= i) {
            const double l = Lq[(int64_t)i * ldq + j];
            acc = fma(l, l, acc);
            double dk = l;
            if (j == i) {
                acc -= log(l * l);
                dk -= 1.0 / l;
            }
            if (apply) g_lq[e] = scale * g_lq[e] - dk;
        } else if (apply) {
            g_lq[e] = 0.0;
        }
    }
    sm[tid] = acc;
    __syncthreads();
    for (int s = nt >> 1; s > 0; s >>= 1) {
        if (tid < s) sm[tid] += sm[tid + s];
        __syncthreads();
    }
    if (tid == 0) out2[0] = 0.5 * (sm[0] - (double)M);
}

__global__ void sum_pairs_kernel(const double* __restrict__ blk, int n, double* __restrict__ out) {
    __shared__ double sm[2][256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += blk[2 * i];
        b += blk[2 * i + 1];
    }
    sm[0][threadIdx.x] = a;
    sm[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + s];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sm[0][0];
        out[1] = sm[1][0];
    }
}

static inline int pad_dims_svgp(int D) {
    int dp = 1;
    while (dp < D) dp <<= 1;
    return dp;
}

#define SVGP_DISPATCH_DP(D, CALL)                                  \
    switch (pad_dims_svgp(D)) {                                    \
        case 1: { constexpr int DP = 1; CALL; } break;             \
        case 2: { constexpr int DP = 2; CALL; } break;             \
        case 4: { constexpr int DP = 4; CALL; } break;             \
        case 8: { constexpr int DP = 8; CALL; } break;             \
        default: { constexpr int DP = 16; CALL; } break;           \
    }

// weights Wt [Na, Nb] against k(A_i, B_j): theta gradient into g_theta (accumulate), first-argument
// gradient into g_rows [Na, D] (accumulate) scaled by zscale
static int cross_grad(gpb_handle* h, const DevKernel& kp, const double* Arows, int64_t Na, const double* Bcols,
                      int64_t Nb, int D, const double* Wt, int64_t ldw, double zscale, double* g_theta,
                      double* g_rows, int accumulate) {
    const int row_tiles = (int)((Na + CG_TILE - 1) / CG_TILE);
    const int col_tiles = (int)((Nb + CG_TILE - 1) / CG_TILE);
    int nsplit = (8 * h->sm_count + row_tiles - 1) / row_tiles;
    if (nsplit > col_tiles) nsplit = col_tiles;
    if (nsplit < 1) nsplit = 1;
    const int dp = pad_dims_svgp(D);
    const size_t tp = (size_t)row_tiles * nsplit * GPB_MAX_PARAMS, zp = (size_t)nsplit * Na * dp;
    double* part = workspace(h, BUF_RED, (tp + zp) * sizeof(double));
    if (!part) return -1;
    double* zpart = part + tp;
    ProfScope prof(h, PROF_SVGP, h->stream);
    dim3 grid((unsigned)row_tiles, (unsigned)nsplit);
    const int shape = h->use_shapes ? match_shape(kp) : SHAPE_NONE;
#define GPB_SHAPE_BODY_                                                                                                  \
    SVGP_DISPATCH_DP(D, (cross_grad_kernel<DP, GPB_SH_FOR(DP)><<<grid, CG_THREADS, 0, h->stream>>>(kp, Arows, (int)Na, Bcols, (int)Nb, \
                                                                                        D, Wt, ldw, zscale, part, zpart)));
    GPB_DISPATCH_SHAPE(shape)
#undef GPB_SHAPE_BODY_
    int rc = check_cuda(h, cudaGetLastError(), "cross_grad_kernel launch");
    if (rc) return rc;
    reduce_rows_kernel<<<kp.n_params, 256, 0, h->stream>>>(part, (int64_t)row_tiles * nsplit, GPB_MAX_PARAMS, 1.0, g_theta,
                                                           accumulate);
    reduce_zbar_kernel<<<(unsigned)((Na * D + 255) / 256), 256, 0, h->stream>>>(zpart, nsplit, (int)Na, dp, D, g_rows,
                                                                                 accumulate);
    h->launches += 3;
    return check_cuda(h, cudaGetLastError(), "cross_grad reduce launch");
}

struct SvgpBuffers {
    double *Kuu, *Wm, *logdiag, *kdiag, *fmean, *fvar, *rvec, *blk, *cpart;
    int* info;
    double *Kuf, *A, *C, *G;  // [M, ldb]
    int64_t ldm, ldb;
    int nch, nblk;
};

static int svgp_alloc(gpb_handle* h, int64_t M, int64_t B, bool need_cg, SvgpBuffers* s) {
    s->ldm = rup(M, 16);
    s->ldb = rup(B, 16);
    const size_t mat = (size_t)rup(M, 128) * s->ldm;
    s->Kuu = workspace(h, BUF_K, mat * sizeof(double));
    s->Wm = workspace(h, BUF_W, mat * sizeof(double));
    s->nch = (int)((M + CS_CH - 1) / CS_CH);
    s->nblk = (int)((B + 255) / 256);
    const size_t vec = (size_t)(rup((M + 127) / 128, 16) + 4 * s->ldb + 2 * (size_t)s->nblk + 3 * (size_t)s->nch * B + 64);
    double* v = workspace(h, BUF_VEC, vec * sizeof(double));
    const size_t big = (size_t)M * s->ldb;
    double* a = workspace(h, BUF_AUX, big * 2 * sizeof(double));
    double* c = need_cg ? workspace(h, BUF_AUX2, big * 2 * sizeof(double)) : workspace(h, BUF_AUX2, big * sizeof(double));
    if (!s->Kuu || !s->Wm || !v || !a || !c) return -1;
    s->logdiag = v;
    s->kdiag = v + rup((M + 127) / 128, 16);
    s->fmean = s->kdiag + s->ldb;
    s->fvar = s->fmean + s->ldb;
    s->rvec = s->fvar + s->ldb;
    s->blk = s->rvec + s->ldb;
    s->cpart = s->blk + 2 * (size_t)s->nblk;
    s->info = reinterpret_cast<int*>(s->cpart + 3 * (size_t)s->nch * B);
    s->Kuf = a;
    s->A = a + big;
    s->C = c;
    s->G = need_cg ? c + big : nullptr;
    return 0;
}

// forward up to the per-column statistics; y may be null (predict)
static int svgp_forward(gpb_handle* h, const DevKernel& kp, const double* d_Z, int64_t M, int D, const double* d_qmu,
                        const double* d_Lq, int64_t ldq, const double* d_X, const double* d_y, int64_t B, double s2,
                        bool keepL, SvgpBuffers& s) {
    int rc;
    if ((rc = launch_assemble(h, kp, d_Z, M, d_Z, M, D, s.Kuu, s.ldm, 1, 1e-6))) return rc;   // default_jitter
    if ((rc = factor_inv(h, s.Kuu, s.ldm, s.Wm, s.ldm, M, s.logdiag, s.info, keepL))) return rc;
    if ((rc = launch_assemble(h, kp, d_Z, M, d_X, B, D, s.Kuf, s.ldb, 0, 0.0))) return rc;
    if ((rc = launch_kdiag(h, kp, d_X, B, D, s.kdiag))) return rc;
    GemmArgs g;
    g.transa = 0; g.transb = 0; g.M = M; g.N = B; g.K = M;          // A = Wm Kuf
    g.A = s.Wm; g.lda = s.ldm; g.B = s.Kuf; g.ldb = s.ldb; g.C = s.A; g.ldc = s.ldb; g.a_lower = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    g = GemmArgs();
    g.transa = 1; g.transb = 0; g.M = M; g.N = B; g.K = M;          // C = Lq^T A
    g.A = d_Lq; g.lda = ldq; g.B = s.A; g.ldb = s.ldb; g.C = s.C; g.ldc = s.ldb; g.a_upper = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    {
        ProfScope prof(h, PROF_SVGP, h->stream);
        dim3 grid((unsigned)((B + 127) / 128), (unsigned)s.nch);
        colstats_partial_kernel<<<grid, 128, 0, h->stream>>>(s.A, s.C, s.ldb, (int)M, (int)B, d_qmu, s.cpart,
                                                             s.cpart + (size_t)s.nch * B, s.cpart + 2 * (size_t)s.nch * B);
        colstats_finish_kernel<<<s.nblk, 256, 0, h->stream>>>(s.cpart, s.cpart + (size_t)s.nch * B,
                                                              s.cpart + 2 * (size_t)s.nch * B, s.nch, (int)B, s.kdiag, d_y, s2,
                                                              s.fmean, s.fvar, s.rvec, s.blk);
        h->launches += 2;
    }
    return check_cuda(h, cudaGetLastError(), "svgp colstats launch");
}

// Shared tail of the SVGP and SGPR adjoints.  On entry s.Kuf holds Kuf_bar = d objective / d k(Z, X)
// [M, B], s.A holds A = Wm k(Z, X), s.Kuu holds Lm (keepL) and s.Wm = Lm^-1; gv = d objective / d k(x_b, x_b).
// Pulls the Cholesky adjoint back to Kuu_bar (the A = Lm^-1 Kuf dependence) and contracts Kuf_bar,
// Kuu_bar and gv with dk/dtheta, dk/dz into g_theta [P] and g_Z [M, D] (overwritten).
// scr: 4 M x ldm doubles (Lm_bar, Q, Sym, T1).
static int sparse_backward_tail(gpb_handle* h, const DevKernel& kp, const double* d_Z, int64_t M, int D, const double* d_X,
                                int64_t B, SvgpBuffers& s, double gv, double* scr, double* g_theta, double* g_Z) {
    int rc;
    GemmArgs g;
    const int P = kp.n_params;
    double* Lbar = scr;
    double* Q = Lbar + (size_t)M * s.ldm;
    double* Sym = Q + (size_t)M * s.ldm;
    double* T1 = Sym + (size_t)M * s.ldm;
    // Lm_bar = -tril(Kuf_bar A^T)
    g = GemmArgs();
    g.transa = 0; g.transb = 1; g.M = M; g.N = M; g.K = B; g.alpha = -1.0;
    g.A = s.Kuf; g.lda = s.ldb; g.B = s.A; g.ldb = s.ldb; g.C = Lbar; g.ldc = s.ldm; g.tri = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    {
        dim3 grid((unsigned)((M + 255) / 256), (unsigned)M);
        zero_strict_upper_kernel<<<grid, 256, 0, h->stream>>>(Lbar, s.ldm, (int)M);
        h->launches += 1;
    }
    // Q = Lm^T Lm_bar (lower tiles) ; Sym = sym-from-lower(Q)
    g = GemmArgs();
    g.transa = 1; g.transb = 0; g.M = M; g.N = M; g.K = M;
    g.A = s.Kuu; g.lda = s.ldm; g.B = Lbar; g.ldb = s.ldm; g.C = Q; g.ldc = s.ldm; g.tri = 1; g.a_upper = 1; g.b_lower = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    {
        dim3 grid((unsigned)((M + 255) / 256), (unsigned)M);
        sym_from_lower_kernel<<<grid, 256, 0, h->stream>>>(Q, s.ldm, Sym, s.ldm, (int)M);
        h->launches += 1;
    }
    // Kuu_bar = 1/2 Wm^T Sym Wm
    g = GemmArgs();
    g.transa = 0; g.transb = 0; g.M = M; g.N = M; g.K = M;
    g.A = Sym; g.lda = s.ldm; g.B = s.Wm; g.ldb = s.ldm; g.C = T1; g.ldc = s.ldm; g.b_lower = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    double* Kuubar = Q;  // Q is dead
    g = GemmArgs();
    g.transa = 1; g.transb = 0; g.M = M; g.N = M; g.K = M; g.alpha = 0.5;
    g.A = s.Wm; g.lda = s.ldm; g.B = T1; g.ldb = s.ldm; g.C = Kuubar; g.ldc = s.ldm; g.a_upper = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    // theta / Z gradients: Kuf_bar against k(Z, X); Kuu_bar against k(Z, Z) (both arguments -> factor 2); kdiag
    if ((rc = cross_grad(h, kp, d_Z, M, d_X, B, D, s.Kuf, s.ldb, 1.0, g_theta, g_Z, 0))) return rc;
    if ((rc = cross_grad(h, kp, d_Z, M, d_Z, M, D, Kuubar, s.ldm, 2.0, g_theta, g_Z, 1))) return rc;
    {
        const int nb = 2 * h->sm_count;
        double* part = workspace(h, BUF_RED, (size_t)nb * GPB_MAX_PARAMS * sizeof(double));
        if (!part) return -1;
        SVGP_DISPATCH_DP(D, (kdiag_grad_kernel<DP><<<nb, 256, 0, h->stream>>>(kp, d_X, (int)B, D, gv, part)));
        reduce_rows_kernel<<<P, 256, 0, h->stream>>>(part, nb, GPB_MAX_PARAMS, 1.0, g_theta, 1);
        h->launches += 2;
    }
    return check_cuda(h, cudaGetLastError(), "sparse backward launches");
}

// q_sqrt must carry zeros above the diagonal inside 128-aligned diagonal blocks (engine convention);
// the host layer passes tril(q_sqrt).
int svgp_data_term(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D,
                   const double* d_qmu, const double* d_Lq, int64_t ldq, const double* d_X, const double* d_y, int64_t B,
                   double* d_flat, int want_grad) {
    if (!h->has_spec) return set_error(h, -3, "svgp: no kernel set");
    if (M <= 0 || B <= 0) return set_error(h, -2, "svgp: empty problem");
    if (!(s2 > 0.0)) return set_error(h, -2, "svgp: noise variance must be > 0");
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "svgp: kernel expects D=%d, got %d", kp.n_dims, D);
    if (want_grad && (kp.n_leaves > GRAD_FAST_LEAVES || kp.has_ard))
        return set_error(h, -4, "svgp gradient supports kernels with <= %d leaves and scalar lengthscales", GRAD_FAST_LEAVES);
    SvgpBuffers s;
    if ((rc = svgp_alloc(h, M, B, want_grad != 0, &s))) return rc;
    if ((rc = svgp_forward(h, kp, d_Z, M, D, d_qmu, d_Lq, ldq, d_X, d_y, B, s2, want_grad != 0, s))) return rc;
    const int P = kp.n_params;
    {
        cudaError_t e = cudaMemsetAsync(d_flat, 0, (size_t)(2 + P + M * D + M + M * M) * sizeof(double), h->stream);
        if (e != cudaSuccess) return check_cuda(h, e, "svgp memset");
    }
    double* g_theta = d_flat + 2;
    double* g_Z = g_theta + P;
    double* g_qmu = g_Z + M * D;
    double* g_Lq = g_qmu + M;
    sum_pairs_kernel<<<1, 256, 0, h->stream>>>(s.blk, s.nblk, d_flat);
    h->launches += 1;
    if (!want_grad) return check_cuda(h, cudaGetLastError(), "svgp sum launch");
    const double gv = -0.5 / s2;
    GemmArgs g;
    // gq_mu = A r
    rowdot_kernel<<<(unsigned)((M + 7) / 8), 256, 0, h->stream>>>(s.A, s.ldb, (int)M, (int)B, s.rvec, g_qmu);
    h->launches += 1;
    // gLq = 2 gv A C^T (lower tiles)
    g = GemmArgs();
    g.transa = 0; g.transb = 1; g.M = M; g.N = M; g.K = B; g.alpha = 2.0 * gv;
    g.A = s.A; g.lda = s.ldb; g.B = s.C; g.ldb = s.ldb; g.C = g_Lq; g.ldc = M; g.tri = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    // G = Lq C ; Abar = q_mu r^T + 2 gv (G - A)
    g = GemmArgs();
    g.transa = 0; g.transb = 0; g.M = M; g.N = B; g.K = M;
    g.A = d_Lq; g.lda = ldq; g.B = s.C; g.ldb = s.ldb; g.C = s.G; g.ldc = s.ldb; g.a_lower = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    {
        dim3 grid((unsigned)((B + 255) / 256), (unsigned)M);
        abar_kernel<<<grid, 256, 0, h->stream>>>(s.G, s.A, s.ldb, (int)M, (int)B, d_qmu, s.rvec, 2.0 * gv);
        h->launches += 1;
    }
    // Kuf_bar = Wm^T Abar  -> into the Kuf buffer (Kuf itself is no longer needed)
    g = GemmArgs();
    g.transa = 1; g.transb = 0; g.M = M; g.N = B; g.K = M;
    g.A = s.Wm; g.lda = s.ldm; g.B = s.G; g.ldb = s.ldb; g.C = s.Kuf; g.ldc = s.ldb; g.a_upper = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    double* scr = workspace(h, BUF_PANEL, (size_t)4 * M * s.ldm * sizeof(double));
    if (!scr) return -1;
    return sparse_backward_tail(h, kp, d_Z, M, D, d_X, B, s, gv, scr, g_theta, g_Z);
}

int svgp_finish(gpb_handle* h, double* d_flat, double scale, const double* d_qmu, const double* d_Lq, int64_t ldq,
                int64_t M, int D, int P, int apply_grad, double* h_elbo, double* h_kl) {
    double* g_theta = d_flat + 2;
    double* g_qmu = g_theta + P + M * D;
    double* g_Lq = g_qmu + M;
    double* v = workspace(h, BUF_DINV, 16 * sizeof(double));
    if (!v) return -1;
    kl_finish_kernel<<<1, 1024, 0, h->stream>>>(d_qmu, d_Lq, ldq, (int)M, scale, d_flat + 1, (int)(1 + P + M * D), g_qmu, g_Lq,
                                                v, apply_grad);
    h->launches += 1;
    int rc = check_cuda(h, cudaGetLastError(), "kl_finish_kernel launch");
    if (rc) return rc;
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp, d_flat, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hp + 1, v, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "svgp finish sync");
    *h_kl = hp[1];
    *h_elbo = scale * hp[0] - hp[1];
    return 0;
}

int svgp_predict_f(gpb_handle* h, const double* theta, const double* d_Z, int64_t M, int D, const double* d_qmu,
                   const double* d_Lq, int64_t ldq, const double* d_Xs, int64_t Ns, double* d_mean, double* d_var) {
    if (!h->has_spec) return set_error(h, -3, "svgp: no kernel set");
    if (Ns <= 0) return 0;
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "svgp: kernel expects D=%d, got %d", kp.n_dims, D);
    int64_t chunk = (int64_t)(1 << 26) / (M > 0 ? M : 1);
    chunk = chunk / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (chunk > Ns) chunk = Ns;
    SvgpBuffers s;
    if ((rc = svgp_alloc(h, M, chunk, false, &s))) return rc;
    for (int64_t s0 = 0; s0 < Ns; s0 += chunk) {
        const int64_t m = (Ns - s0 < chunk) ? (Ns - s0) : chunk;
        if ((rc = svgp_forward(h, kp, d_Z, M, D, d_qmu, d_Lq, ldq, d_Xs + s0 * D, nullptr, m, 1.0, false, s))) return rc;
        cudaError_t e = cudaMemcpyAsync(d_mean + s0, s.fmean, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_var + s0, s.fvar, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) return check_cuda(h, e, "svgp predict copy");
    }
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp + 64, s.info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "svgp predict sync");
    const int info = *reinterpret_cast<int*>(hp + 64);
    if (info > 0) {
        set_error(h, info, "Cholesky decomposition was not successful: Kuu pivot %d", info);
        return info;
    }
    return 0;
}


// ====================================================================================================
// SGPR -- Titsias' collapsed sparse bound (gpflow/models/sgpr.py SGPR.elbo / predict_f), the model the
// reference builds at test_scripts/SVGP.py:393-399 and trains with Scipy.
// Shares Kuu / Kuf assembly, the blocked factorisation and the adjoint tail with SVGP.
//
// Forward (s = noise variance, err = y - m(X), jitter = 1e-6):
//   Lm = chol(k(Z,Z) + jitter I), Wm = Lm^-1          V = Wm k(Z,X)                     [M,N]
//   Bm = I + V V^T / s, LB = chol(Bm), WB = LB^-1     c = WB V err / s,  beta = WB^T c  (= Bm^-1 V err / s)
//   elbo = -N/2 log 2pi - sum log LB_ii - N/2 log s - (sum_n kdiag_n - tr V V^T)/(2s) - |err|^2/(2s) + |c|^2/2
// Adjoint (hand-derived; replaces TF autodiff through two Choleskys):
//   H = I - Bm^-1 - beta beta^T                        V_bar = (H V + beta err^T) / s
//   Kuf_bar = Wm^T V_bar ; Lm_bar, Kuu_bar and the theta / Z contractions: sparse_backward_tail
//   d/dkdiag_n = -1/(2s) ; err_bar = (V^T beta - err) / s
//   d/ds = [(M - tr Bm^-1)/2 - N/2 - (|c|^2 + |beta|^2)/2]/s + [sum kdiag - tr V V^T + |err|^2]/(2 s^2)
// Three M x M x N products (V, V V^T lower, H-weighted V) plus the tail's one; M^3 work is minor.

// out[0] = trace(B) before the shift (fixed order, one block); then B_ii += 1
__global__ void sgpr_shift_trace_kernel(double* __restrict__ B, int64_t ld, int M, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) {
        const double d = B[(int64_t)i * ld + i];
        s += d;
        B[(int64_t)i * ld + i] = d + 1.0;
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// out[0] = sum a_i^2 (a may be null -> 0), out[1] = sum b_i (b may be null -> 0); one block, fixed order
__global__ void sgpr_sums_kernel(const double* __restrict__ a, int64_t na, const double* __restrict__ b, int64_t nb,
                                 double* __restrict__ out) {
    __shared__ double sm[2][1024];
    double sa = 0.0, sb = 0.0;
    if (a) for (int64_t i = threadIdx.x; i < na; i += 1024) sa = fma(a[i], a[i], sa);
    if (b) for (int64_t i = threadIdx.x; i < nb; i += 1024) sb += b[i];
    sm[0][threadIdx.x] = sa;
    sm[1][threadIdx.x] = sb;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + k];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sm[0][0]; out[1] = sm[1][0]; }
}

// out[0] = trace of a lower-stored matrix (one block, fixed order)
__global__ void sgpr_trace_kernel(const double* __restrict__ A, int64_t ld, int M, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) s += A[(int64_t)i * ld + i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

__global__ void scale_vec_kernel(double* __restrict__ v, int n, double a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= a;
}

// H = I - Binv - beta beta^T (full symmetric, from the lower-stored Binv)
__global__ void sgpr_h_kernel(const double* __restrict__ Binv, int64_t ldb, int M, const double* __restrict__ beta,
                              double* __restrict__ H, int64_t ldh) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= M) return;
    const double bi = (i >= j) ? Binv[(int64_t)i * ldb + j] : Binv[(int64_t)j * ldb + i];
    H[(int64_t)i * ldh + j] = ((i == j) ? 1.0 : 0.0) - bi - beta[i] * beta[j];
}

// G[m][b] += scale * w[m] * e[b]
__global__ void rank1_add_kernel(double* __restrict__ G, int64_t ld, int M, int64_t N, const double* __restrict__ w,
                                 const double* __restrict__ e, double scale) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (b >= N) return;
    const int64_t o = (int64_t)m * ld + b;
    G[o] = fma(scale * w[m], e[b], G[o]);
}

// out[b] = (sum_m V[m][b] beta[m] - err[b]) / s
__global__ void sgpr_errbar_kernel(const double* __restrict__ V, int64_t ld, int M, int64_t N, const double* __restrict__ beta,
                                   const double* __restrict__ err, double inv_s, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    double s = 0.0;
    for (int m = 0; m < M; ++m) s = fma(V[(int64_t)m * ld + b], beta[m], s);
    out[b] = (s - err[b]) * inv_s;
}

struct SgprState {
    SvgpBuffers s;       // Kuu (-> Lm), Wm, Kuf, A (= V), ldm, ldb
    double *Bm, *WB, *mm2, *mm3, *scr;   // M x ldm each; scr = 4 M x ldm for the tail
    double *u, *c, *beta, *wbeta, *logdiagM, *logdiagB, *kdiag, *scal, *flat;
    int *info;           // [0] Kuu, [1] Bm
};

static int sgpr_forward(gpb_handle* h, const DevKernel& kp, double s2, const double* d_Z, int64_t M, int D, const double* d_X,
                        const double* d_err, int64_t N, bool want_grad, int P, SgprState* st) {
    SvgpBuffers& s = st->s;
    s.ldm = rup(M, 16);
    s.ldb = rup(N, 16);
    const size_t mat = (size_t)rup(M, 128) * s.ldm;
    const size_t big = (size_t)M * s.ldb;
    s.Kuu = workspace(h, BUF_K, mat * sizeof(double));
    s.Wm = workspace(h, BUF_W, mat * sizeof(double));
    s.Kuf = workspace(h, BUF_AUX, big * sizeof(double));
    s.A = workspace(h, BUF_AUX2, big * sizeof(double));
    const int64_t nbk = rup((M + 127) / 128, 16);
    const size_t small = (size_t)(4 * s.ldm + 2 * nbk + s.ldb + 16 + 2 + P + M * D + 16);
    double* v = workspace(h, BUF_VEC, small * sizeof(double));
    if (!s.Kuu || !s.Wm || !s.Kuf || !s.A || !v) return -1;
    st->u = v;
    st->c = st->u + s.ldm;
    st->beta = st->c + s.ldm;
    st->wbeta = st->beta + s.ldm;
    st->logdiagM = st->wbeta + s.ldm;
    st->logdiagB = st->logdiagM + nbk;
    st->kdiag = st->logdiagB + nbk;
    st->scal = st->kdiag + s.ldb;    // [0] tr(VV^T)/s [1] |err|^2 [2] sum kdiag [3] |c|^2 [4] sum log LB [5] tr Bm^-1 [6] |beta|^2
    st->flat = st->scal + 16;        // [2 + P + M D]
    st->info = reinterpret_cast<int*>(st->flat + 2 + P + M * D);
    int rc;
    if ((rc = launch_assemble(h, kp, d_Z, M, d_Z, M, D, s.Kuu, s.ldm, 1, 1e-6))) return rc;   // default_jitter
    if ((rc = factor_inv(h, s.Kuu, s.ldm, s.Wm, s.ldm, M, st->logdiagM, st->info, want_grad))) return rc;
    // persistent M x M scratch is taken only now: factor_inv(keepL) borrows BUF_PANEL while it runs
    double* mm = workspace(h, BUF_PANEL, (size_t)(4 * mat + 4 * (size_t)M * s.ldm) * sizeof(double));
    if (!mm) return -1;
    st->Bm = mm;
    st->WB = mm + mat;
    st->mm2 = mm + 2 * mat;
    st->mm3 = mm + 3 * mat;
    st->scr = mm + 4 * mat;
    if ((rc = launch_assemble(h, kp, d_Z, M, d_X, N, D, s.Kuf, s.ldb, 0, 0.0))) return rc;
    if ((rc = launch_kdiag(h, kp, d_X, N, D, st->kdiag))) return rc;
    GemmArgs g;
    g.transa = 0; g.transb = 0; g.M = M; g.N = N; g.K = M;          // V = Wm Kuf
    g.A = s.Wm; g.lda = s.ldm; g.B = s.Kuf; g.ldb = s.ldb; g.C = s.A; g.ldc = s.ldb; g.a_lower = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    g = GemmArgs();
    g.transa = 0; g.transb = 1; g.M = M; g.N = M; g.K = N; g.alpha = 1.0 / s2;   // Bm - I = V V^T / s (lower tiles)
    g.A = s.A; g.lda = s.ldb; g.B = s.A; g.ldb = s.ldb; g.C = st->Bm; g.ldc = s.ldm; g.tri = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    sgpr_shift_trace_kernel<<<1, 256, 0, h->stream>>>(st->Bm, s.ldm, (int)M, st->scal + 0);
    sgpr_sums_kernel<<<1, 1024, 0, h->stream>>>(d_err, N, st->kdiag, N, st->scal + 1);
    h->launches += 2;
    if ((rc = factor_inv(h, st->Bm, s.ldm, st->WB, s.ldm, M, st->logdiagB, st->info + 1, false))) return rc;
    rowdot_kernel<<<(unsigned)((M + 7) / 8), 256, 0, h->stream>>>(s.A, s.ldb, (int)M, (int)N, d_err, st->u);   // u = V err
    h->launches += 1;
    if ((rc = trmv_lower(h, st->WB, s.ldm, M, st->u, st->c))) return rc;
    scale_vec_kernel<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(st->c, (int)M, 1.0 / s2);              // c = WB u / s
    h->launches += 1;
    if ((rc = trmv_lower_T(h, st->WB, s.ldm, M, st->c, st->beta))) return rc;                                  // beta = WB^T c
    if ((rc = quad_logdet(h, st->c, M, st->logdiagB, st->scal + 3))) return rc;                               // |c|^2, sum log LB
    return check_cuda(h, cudaGetLastError(), "sgpr forward launches");
}

static int sgpr_check_info(gpb_handle* h, const int* info) {
    if (info[0] > 0) {
        set_error(h, info[0], "Cholesky decomposition was not successful: Kuu pivot %d", info[0]);
        return info[0];
    }
    if (info[1] > 0) {
        set_error(h, info[1], "Cholesky decomposition was not successful: I + A A^T pivot %d", info[1]);
        return info[1];
    }
    return 0;
}

// h_out [2 + P + M*D] = elbo, d elbo/d noise, d elbo/d theta (constrained), d elbo/d Z (row-major);
// d_errbar [N] (optional) = d elbo / d err (the host layer chains it through the mean function).
int sgpr_elbo(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D, const double* d_X,
              const double* d_err, int64_t N, int want_grad, double* h_out, double* d_errbar) {
    if (!h->has_spec) return set_error(h, -3, "sgpr: no kernel set");
    if (M <= 0 || N <= 0) return set_error(h, -2, "sgpr: empty problem");
    if (!(s2 > 0.0)) return set_error(h, -2, "sgpr: noise variance must be > 0");
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "sgpr: kernel expects D=%d, got %d", kp.n_dims, D);
    if (want_grad && (kp.n_leaves > GRAD_FAST_LEAVES || kp.has_ard))
        return set_error(h, -4, "sgpr gradient supports kernels with <= %d leaves and scalar lengthscales", GRAD_FAST_LEAVES);
    const int P = kp.n_params;
    SgprState st;
    if ((rc = sgpr_forward(h, kp, s2, d_Z, M, D, d_X, d_err, N, want_grad != 0, P, &st))) return rc;
    SvgpBuffers& s = st.s;
    const size_t nflat = (size_t)(2 + P + M * D);
    if (want_grad) {
        double* g_theta = st.flat + 2;
        double* g_Z = g_theta + P;
        // Bm^-1 = WB^T WB (lower tiles) ; H = I - Bm^-1 - beta beta^T
        if ((rc = lauum_lower(h, st.WB, M, s.ldm, st.mm2, s.ldm))) return rc;
        sgpr_trace_kernel<<<1, 256, 0, h->stream>>>(st.mm2, s.ldm, (int)M, st.scal + 5);
        sgpr_sums_kernel<<<1, 1024, 0, h->stream>>>(st.beta, M, nullptr, 0, st.scal + 6);
        {
            dim3 grid((unsigned)((M + 255) / 256), (unsigned)M);
            sgpr_h_kernel<<<grid, 256, 0, h->stream>>>(st.mm2, s.ldm, (int)M, st.beta, st.mm3, s.ldm);
        }
        h->launches += 3;
        // G1 = Wm^T H  (into the dead Bm buffer) ; Kuf_bar = G1 V / s + (Wm^T beta) err^T / s
        GemmArgs g;
        g.transa = 1; g.transb = 0; g.M = M; g.N = M; g.K = M;
        g.A = s.Wm; g.lda = s.ldm; g.B = st.mm3; g.ldb = s.ldm; g.C = st.Bm; g.ldc = s.ldm; g.a_upper = 1;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        g = GemmArgs();
        g.transa = 0; g.transb = 0; g.M = M; g.N = N; g.K = M; g.alpha = 1.0 / s2;
        g.A = st.Bm; g.lda = s.ldm; g.B = s.A; g.ldb = s.ldb; g.C = s.Kuf; g.ldc = s.ldb;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        if ((rc = trmv_lower_T(h, s.Wm, s.ldm, M, st.beta, st.wbeta))) return rc;
        {
            dim3 grid((unsigned)((N + 255) / 256), (unsigned)M);
            rank1_add_kernel<<<grid, 256, 0, h->stream>>>(s.Kuf, s.ldb, (int)M, N, st.wbeta, d_err, 1.0 / s2);
            h->launches += 1;
        }
        if (d_errbar) {
            sgpr_errbar_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(s.A, s.ldb, (int)M, N, st.beta, d_err,
                                                                                   1.0 / s2, d_errbar);
            h->launches += 1;
        }
        if ((rc = sparse_backward_tail(h, kp, d_Z, M, D, d_X, N, s, -0.5 / s2, st.scr, g_theta, g_Z))) return rc;
    }
    // one D2H of the scalars, the gradient record and the two pivot flags; the only sync of the evaluation
    double* hp = pinned(h, (nflat + 16 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp, st.scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && want_grad)
        e = cudaMemcpyAsync(hp + 16, st.flat, nflat * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hp + 16 + nflat, st.info, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "sgpr result copy/sync");
    if ((rc = sgpr_check_info(h, reinterpret_cast<int*>(hp + 16 + nflat)))) return rc;
    const double trVV = s2 * hp[0], ee = hp[1], sk = hp[2], cc = hp[3], hl = hp[4];
    h_out[0] = -0.5 * (double)N * log(2.0 * M_PI) - hl - 0.5 * (double)N * log(s2) - 0.5 * (sk - trVV) / s2 - 0.5 * ee / s2 +
               0.5 * cc;
    if (want_grad) {
        const double trBinv = hp[5], bb = hp[6];
        h_out[1] = (0.5 * ((double)M - trBinv) - 0.5 * (double)N - 0.5 * (cc + bb)) / s2 + 0.5 * (sk - trVV + ee) / (s2 * s2);
        for (size_t i = 2; i < nflat; ++i) h_out[i] = hp[16 + i];
    }
    return 0;
}

// SGPR.predict_f(Xnew, full_cov=False): mean = tmp2^T c, var = k** + |tmp2|^2 - |tmp1|^2 with
// tmp1 = Wm k(Z, Xnew), tmp2 = WB tmp1 (mean excludes mean_function(Xnew)).
int sgpr_predict_f(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D, const double* d_X,
                   const double* d_err, int64_t N, const double* d_Xs, int64_t Ns, double* d_mean, double* d_var) {
    if (!h->has_spec) return set_error(h, -3, "sgpr: no kernel set");
    if (M <= 0 || N <= 0) return set_error(h, -2, "sgpr: empty problem");
    if (!(s2 > 0.0)) return set_error(h, -2, "sgpr: noise variance must be > 0");
    if (Ns <= 0) return 0;
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "sgpr: kernel expects D=%d, got %d", kp.n_dims, D);
    SgprState st;
    if ((rc = sgpr_forward(h, kp, s2, d_Z, M, D, d_X, d_err, N, false, kp.n_params, &st))) return rc;
    SvgpBuffers& s = st.s;
    // the [M, N] training buffers are dead now: reuse them for the test chunks
    int64_t chunk = (int64_t)(1 << 26) / M / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (chunk > Ns) chunk = rup(Ns, 16);
    const int64_t ldc = rup(chunk, 16);
    double* Kus = workspace(h, BUF_AUX, (size_t)M * ldc * sizeof(double));
    double* t1 = workspace(h, BUF_AUX2, (size_t)M * ldc * sizeof(double));
    const int nch = (int)((M + CS_CH - 1) / CS_CH);
    double* t2 = workspace(h, BUF_RED, ((size_t)M * ldc + 3 * (size_t)nch * chunk + 3 * (size_t)ldc) * sizeof(double));
    if (!Kus || !t1 || !t2) return -1;
    double* cpart = t2 + (size_t)M * ldc;
    double* kd = cpart + 3 * (size_t)nch * chunk;
    double* fm = kd + ldc;
    double* fv = fm + ldc;
    for (int64_t s0 = 0; s0 < Ns; s0 += chunk) {
        const int64_t m = (Ns - s0 < chunk) ? (Ns - s0) : chunk;
        const double* Xs = d_Xs + s0 * D;
        if ((rc = launch_assemble(h, kp, d_Z, M, Xs, m, D, Kus, ldc, 0, 0.0))) return rc;
        if ((rc = launch_kdiag(h, kp, Xs, m, D, kd))) return rc;
        GemmArgs g;
        g.transa = 0; g.transb = 0; g.M = M; g.N = m; g.K = M;
        g.A = s.Wm; g.lda = s.ldm; g.B = Kus; g.ldb = ldc; g.C = t1; g.ldc = ldc; g.a_lower = 1;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        g.A = st.WB; g.B = t1; g.C = t2;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        dim3 grid((unsigned)((m + 127) / 128), (unsigned)nch);
        colstats_partial_kernel<<<grid, 128, 0, h->stream>>>(t1, t2, ldc, (int)M, (int)m, st.beta, cpart, cpart + (size_t)nch * m,
                                                             cpart + 2 * (size_t)nch * m);
        colstats_finish_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(
            cpart, cpart + (size_t)nch * m, cpart + 2 * (size_t)nch * m, nch, (int)m, kd, nullptr, 1.0, fm, fv, nullptr, nullptr);
        h->launches += 2;
        cudaError_t e = cudaMemcpyAsync(d_mean + s0, fm, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_var + s0, fv, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) return check_cuda(h, e, "sgpr predict copy");
    }
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp + 64, st.info, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "sgpr predict sync");
    return sgpr_check_info(h, reinterpret_cast<int*>(hp + 64));
}

}  // namespace gpb

namespace gpb {

// Adam on a device-resident parameter block (the SVGP variational parameters and inducing points
// have identity transforms, so they are updated in place on the device; the handful of
// constrained hyper-parameters are stepped by the host layer).  sign = +1 ascends (ELBO).
__global__ void adam_step_kernel(double* __restrict__ x, const double* __restrict__ g, double* __restrict__ m,
                                 double* __restrict__ v, int64_t n, double lr, double b1, double b2, double eps,
                                 double c1, double c2, double sign) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double gi = g[i];
    const double mi = b1 * m[i] + (1.0 - b1) * gi;
    const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    x[i] += sign * lr * (mi / c1) / (sqrt(vi / c2) + eps);
}

int adam_step(gpb_handle* h, double* d_x, const double* d_g, double* d_m, double* d_v, int64_t n, double lr, double b1,
              double b2, double eps, int64_t step, double sign) {
    if (n <= 0) return 0;
    const double c1 = 1.0 - pow(b1, (double)step), c2 = 1.0 - pow(b2, (double)step);
    adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_x, d_g, d_m, d_v, n, lr, b1, b2, eps, c1, c2, sign);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "adam_step_kernel launch");
}

}  // namespace gpb

extern "C" int gpb_adam_step(gpb_handle* h, double* d_x, const double* d_g, double* d_m, double* d_v, int64_t n, double lr,
                             double beta1, double beta2, double eps, int64_t step, int maximize) {
    GPB_ENTER(h);
    if (!d_x || !d_g || !d_m || !d_v || step < 1) return gpb::set_error(h, -2, "adam_step: bad arguments");
    return gpb::adam_step(h, d_x, d_g, d_m, d_v, n, lr, beta1, beta2, eps, step, maximize ? 1.0 : -1.0);
}
