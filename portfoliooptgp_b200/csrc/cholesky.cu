// cholesky.cu -- blocked fp64 Cholesky, triangular inverse and (L L^T)^-1 for one large exact GP
// (north_star subsystem 2).  Replaces tf.linalg.cholesky / tf.linalg.triangular_solve and the
// O(N^3) passes of TF's CholeskyGrad that GPflow's GPR objective triggers (SURVEY.md 2.1 K2, K3,
// K5; reference call sites GPR/model_trainer.py:18-19, Multi-Input_GPR/models/model_trainer.py:21).
//
// Algorithm (all matrices row-major, lower triangular, blocks aligned to NB = 128):
//   factor_inv(A, W):  recursive 2x2 splitting
//       [A11    ]      L11, W11 = factor_inv(A11)                      (recursion; leaf = one CTA)
//       [A21 A22]      T   = A21 W11^T            (= L21)              DMMA GEMM, k-range limited
//                      A22 -= T T^T               (trailing SYRK)      DMMA GEMM, lower tiles only
//                      L22, W22 = factor_inv(A22)
//                      U   = T W11 ; W21 = -W22 U (inverse of the factor, needed for K^-1)
//   lauum: K^-1 = W^T W (lower tiles).
// Every O(N^3) flop runs in dgemm.cu on the FP64 tensor pipe; the only non-GEMM work is the
// 128x128 leaf (block_chol.cuh), one CTA, latency-bound.
// Flops: factor N^3/3 + inverse N^3/3 + lauum N^3/3 = N^3.
#include "block_chol.cuh"
#include "engine.cuh"

namespace gpb {

constexpr int NB = 128;
constexpr int LEAF_THREADS = 512;
constexpr size_t LEAF_SMEM = (size_t)(NB * SLD + 64 * TLD + DINV_DOUBLES + 16) * sizeof(double);

// One CTA: L = chol(A_blk) in place, W_blk = L^-1, logdiag[blk] = sum log L_ii, info = first bad pivot.
// The block lives in shared memory (stride SLD); factorisation and inverse run on DMMA tiles
// (block_chol.cuh).  Blocks narrower than 128 are padded to a multiple of 8 with an identity.
__global__ void __launch_bounds__(LEAF_THREADS)
leaf_potrf_inv_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ W, int64_t ldw, int n, int offset,
                      double* __restrict__ logdiag, int* __restrict__ info, int store_L) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* T = sm + NB * SLD;
    double* dinv = T + 64 * TLD;
    int* fail = reinterpret_cast<int*>(dinv + DINV_DOUBLES);
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int np = (n + 7) & ~7;
    // 128 x 128 block, 512 threads: 32 elements per thread, all loads in flight before the stores
    {
        double v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int idx = tid + u * LEAF_THREADS;
            const int i = idx >> 7, j = idx & (NB - 1);
            v[u] = (i < n && j <= i) ? A[(int64_t)i * lda + j] : ((i >= n && i == j) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int idx = tid + u * LEAF_THREADS;
            const int i = idx >> 7, j = idx & (NB - 1);
            if (i < np && j < np) S[i * SLD + j] = v[u];
        }
    }
    __syncthreads();
    block_potrf_lower(S, np, fail, dinv);
    if (tid == 0 && *fail != 0) atomicCAS(info, 0, offset + *fail);
    // L itself is only needed by callers that keep the factor (gpb_potrf, SVGP adjoint); the LML / K^-1
    // pipeline consumes W and the log-diagonal only, so the 128 KB store is skipped there
    if (store_L) {
#pragma unroll 8
        for (int idx = tid; idx < n * NB; idx += LEAF_THREADS) {
            const int i = idx >> 7, j = idx & (NB - 1);
            if (j < n) A[(int64_t)i * lda + j] = S[i * SLD + j];
        }
    }
    // sum of log-diagonal, fixed order: warp 0
    if (tid < 32) {
        double s = 0.0;
        for (int i = tid; i < n; i += 32) s += log(S[i * SLD + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (tid == 0) logdiag[offset / NB] = s;
    }
    __syncthreads();
    block_trtri_lower_inplace(S, np, T, dinv);
#pragma unroll 8
    for (int idx = tid; idx < n * NB; idx += LEAF_THREADS) {
        const int i = idx >> 7, j = idx & (NB - 1);
        if (j < n) W[(int64_t)i * ldw + j] = S[i * SLD + j];
    }
}

static int leaf(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int n, int offset, double* logdiag,
                int* info, bool store_L) {
    static bool attr_set[GPB_MAX_DEVICES] = {};
    {
        cudaError_t e = ensure_dyn_smem(attr_set, h->device, leaf_potrf_inv_kernel, LEAF_SMEM);
        if (e != cudaSuccess) return check_cuda(h, e, "leaf cudaFuncSetAttribute");
    }
    ProfScope prof(h, PROF_LEAF, h->stream);
    if (h->use_pdl) {
        cudaError_t le = launch_pdl(leaf_potrf_inv_kernel, dim3(1), dim3(LEAF_THREADS), LEAF_SMEM, h->stream,
                                    A + (int64_t)offset * lda + offset, lda, W + (int64_t)offset * ldw + offset, ldw, n,
                                    offset, logdiag, info, store_L ? 1 : 0);
        if (le != cudaSuccess) return check_cuda(h, le, "leaf launch (PDL)");
    } else {
        leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A + (int64_t)offset * lda + offset, lda,
                                                                         W + (int64_t)offset * ldw + offset, ldw, n, offset,
                                                                         logdiag, info, store_L ? 1 : 0);
    }
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "leaf_potrf_inv_kernel launch");
}

// A, W: full matrices; factor the diagonal block [o, o+n).  keepL: additionally leave L21 in A21
// (needs U scratch of n2 x n1 doubles).
static int factor_inv_rec(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int o, int n, double* logdiag,
                          int* info, bool keepL, double* scratchU, int depth) {
    if (n <= NB) return leaf(h, A, lda, W, ldw, n, o, logdiag, info, keepL);
    const int n1 = ((n / 2 + NB - 1) / NB) * NB, n2 = n - n1, o2 = o + n1;
    int rc = factor_inv_rec(h, A, lda, W, ldw, o, n1, logdiag, info, keepL, scratchU, depth + 1);
    if (rc) return rc;
    double* A21 = A + (int64_t)o2 * lda + o;
    double* A22 = A + (int64_t)o2 * lda + o2;
    double* W11 = W + (int64_t)o * ldw + o;
    double* W21 = W + (int64_t)o2 * ldw + o;
    double* W22 = W + (int64_t)o2 * ldw + o2;
    GemmArgs g;
    // T = A21 W11^T  -> W21 region
    g = GemmArgs();
    g.transa = 0; g.transb = 1; g.M = n2; g.N = n1; g.K = n1;
    g.A = A21; g.lda = lda; g.B = W11; g.ldb = ldw; g.C = W21; g.ldc = ldw; g.b_upper = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    // fork point: everything that only needs T (and W11) may start now
    const bool fork = h->fork_streams && (depth < gpb_handle::MAX_DEPTH) && h->side[depth] && n2 >= NB;
    cudaStream_t us = fork ? h->side[depth] : h->stream;
    if (fork) {
        cudaError_t e = cudaEventRecord(h->ev_fork[depth], h->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(us, h->ev_fork[depth], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "fork U stream");
    }
    // A22 -= T T^T (lower tiles)
    g = GemmArgs();
    g.transa = 0; g.transb = 1; g.M = n2; g.N = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0;
    g.A = W21; g.lda = ldw; g.B = W21; g.ldb = ldw; g.C = A22; g.ldc = lda; g.tri = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    double* U = A21;
    int64_t ldu = lda;
    if (keepL) {
        // keep L21: copy T into A21, route U through scratch
        cudaError_t e = cudaMemcpy2DAsync(A21, lda * sizeof(double), W21, ldw * sizeof(double), (size_t)n1 * sizeof(double),
                                          (size_t)n2, cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) return check_cuda(h, e, "copy L21");
        U = scratchU;
        ldu = n1;
    }
    // U = T W11 (W11 lower: k >= j).  Off the critical path until W21 below: forked onto this depth's
    // side stream so that it fills the SMs left idle by the latency-bound bottom of the A22 subtree.
    g = GemmArgs();
    g.transa = 0; g.transb = 0; g.M = n2; g.N = n1; g.K = n1;
    g.A = W21; g.lda = ldw; g.B = W11; g.ldb = ldw; g.C = U; g.ldc = ldu; g.b_lower = 1;
    if ((rc = launch_gemm(h, g, us))) return rc;
    if (fork) {
        cudaError_t e = cudaEventRecord(h->ev_join[depth], us);
        if (e != cudaSuccess) return check_cuda(h, e, "record U join");
    }
    if ((rc = factor_inv_rec(h, A, lda, W, ldw, o2, n2, logdiag, info, keepL, keepL ? scratchU + (int64_t)n2 * n1 : scratchU,
                             depth + 1)))
        return rc;
    if (fork) {
        cudaError_t e = cudaStreamWaitEvent(h->stream, h->ev_join[depth], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "join U stream");
    }
    // W21 = -W22 U (W22 lower: k <= i)
    g = GemmArgs();
    g.transa = 0; g.transb = 0; g.M = n2; g.N = n1; g.K = n2; g.alpha = -1.0;
    g.A = W22; g.lda = ldw; g.B = U; g.ldb = ldu; g.C = W21; g.ldc = ldw; g.a_lower = 1;
    return launch_gemm(h, g, h->stream);
}

// scratch doubles needed by keepL at size n (sum over the recursion's live U blocks)
static size_t keepL_scratch(int n) {
    if (n <= NB) return 0;
    const int n1 = ((n / 2 + NB - 1) / NB) * NB, n2 = n - n1;
    size_t below = keepL_scratch(n1);
    size_t right = (size_t)n2 * n1 + keepL_scratch(n2);
    return below > right ? below : right;
}

int factor_inv(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int64_t N, double* logdiag, int* d_info,
               bool keepL) {
    double* scratch = nullptr;
    if (keepL) {
        size_t need = keepL_scratch((int)N);
        if (need) {
            scratch = workspace(h, BUF_PANEL, need * sizeof(double));
            if (!scratch) return -1;
        }
    }
    cudaError_t e = cudaMemsetAsync(d_info, 0, sizeof(int), h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "memset info");
    return factor_inv_rec(h, A, lda, W, ldw, 0, (int)N, logdiag, d_info, keepL, scratch, 0);
}

int lauum_lower(gpb_handle* h, const double* d_W, int64_t N, int64_t ldw, double* d_Out, int64_t ldo) {
    GemmArgs g;
    g.transa = 1; g.transb = 0; g.M = N; g.N = N; g.K = N;
    g.A = d_W; g.lda = ldw; g.B = d_W; g.ldb = ldw; g.C = d_Out; g.ldc = ldo; g.tri = 1; g.a_upper = 1;
    return launch_gemm(h, g, h->stream);
}

// ---- triangular matrix-vector products --------------------------------------------------------------
// a = W y (W lower): one warp per row.
__global__ void trmv_lower_kernel(const double* __restrict__ W, int64_t ldw, int n, const double* __restrict__ y,
                                  double* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* w = W + (int64_t)row * ldw;
    double s = 0.0;
    for (int j = lane; j <= row; j += 32) s = fma(w[j], y[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

// partial[c][j] = sum_{i in chunk c, i >= j} W[i][j] a[i]; chunks of CH rows; thread per column.
// (32-row chunks: at N = 1000 the 128-row version was one dependent load-FMA chain of 128 steps on 64
// CTAs and took 56 us; short chunks give the grid enough CTAs and the unrolled loads overlap.)
constexpr int TRMVT_CH = 32;
__global__ void trmvT_partial_kernel(const double* __restrict__ W, int64_t ldw, int n, const double* __restrict__ a,
                                     double* __restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= n) return;
    const int i0 = c * TRMVT_CH, i1 = min(n, i0 + TRMVT_CH);
    double s = 0.0;
    if (j <= i0 && i1 - i0 == TRMVT_CH) {
        // full chunk below the diagonal: all loads issued before the (fixed-order) FMA chain
        double w[TRMVT_CH];
#pragma unroll
        for (int k = 0; k < TRMVT_CH; ++k) w[k] = W[(int64_t)(i0 + k) * ldw + j];
#pragma unroll
        for (int k = 0; k < TRMVT_CH; ++k) s = fma(w[k], a[i0 + k], s);
    } else {
        for (int i = max(i0, j); i < i1; ++i) s = fma(W[(int64_t)i * ldw + j], a[i], s);
    }
    partial[(int64_t)c * n + j] = s;
}
__global__ void colsum_partials_kernel(const double* __restrict__ partial, int nchunks, int n, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(int64_t)c * n + j];
    out[j] = s;
}

int trmv_lower(gpb_handle* h, const double* W, int64_t ldw, int64_t n, const double* y, double* out) {
    const int warps = 8;
    ProfScope prof(h, PROF_VEC, h->stream);
    trmv_lower_kernel<<<(unsigned)((n + warps - 1) / warps), warps * 32, 0, h->stream>>>(W, ldw, (int)n, y, out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "trmv_lower_kernel launch");
}

int trmv_lower_T(gpb_handle* h, const double* W, int64_t ldw, int64_t n, const double* a, double* out) {
    const int nch = (int)((n + TRMVT_CH - 1) / TRMVT_CH);
    double* partial = workspace(h, BUF_RED, (size_t)nch * n * sizeof(double));
    if (!partial) return -1;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)nch);
    ProfScope prof(h, PROF_VEC, h->stream);
    trmvT_partial_kernel<<<grid, 128, 0, h->stream>>>(W, ldw, (int)n, a, partial);
    colsum_partials_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(partial, nch, (int)n, out);
    h->launches += 2;
    return check_cuda(h, cudaGetLastError(), "trmvT kernels launch");
}

// out[0] = sum v_i^2 ; out[1] = sum_b logdiag[b]   (single block, fixed order)
__global__ void quad_logdet_kernel(const double* __restrict__ v, int n, const double* __restrict__ logdiag, int nb,
                                   double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s = fma(v[i], v[i], s);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sm[0];
        double l = 0.0;
        for (int b = 0; b < nb; ++b) l += logdiag[b];
        out[1] = l;
    }
}

int quad_logdet(gpb_handle* h, const double* v, int64_t n, const double* logdiag, double* out) {
    quad_logdet_kernel<<<1, 256, 0, h->stream>>>(v, (int)n, logdiag, (int)((n + NB - 1) / NB), out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "quad_logdet_kernel launch");
}

// Column reductions over a dense [n, m] block: ss[j] = sum_i A[i][j]^2, dot[j] = sum_i A[i][j] a[i].
constexpr int COLRED_CH = 256;
__global__ void colred_partial_kernel(const double* __restrict__ A, int64_t lda, int n, int m, const double* __restrict__ a,
                                      double* __restrict__ p_ss, double* __restrict__ p_dot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= m) return;
    const int i0 = c * COLRED_CH, i1 = min(n, i0 + COLRED_CH);
    double ss = 0.0, dt = 0.0;
    for (int i = i0; i < i1; ++i) {
        const double v = A[(int64_t)i * lda + j];
        ss = fma(v, v, ss);
        dt = fma(v, a[i], dt);
    }
    p_ss[(int64_t)c * m + j] = ss;
    p_dot[(int64_t)c * m + j] = dt;
}
__global__ void predict_finish_kernel(const double* __restrict__ p_ss, const double* __restrict__ p_dot, int nchunks,
                                      int m, const double* __restrict__ kdiag, double* __restrict__ mean,
                                      double* __restrict__ var) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double ss = 0.0, dt = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        ss += p_ss[(int64_t)c * m + j];
        dt += p_dot[(int64_t)c * m + j];
    }
    mean[j] = dt;
    var[j] = kdiag[j] - ss;
}

int predict_colreduce(gpb_handle* h, const double* A, int64_t lda, int64_t n, int64_t m, const double* a,
                      const double* kdiag, double* mean, double* var) {
    const int nch = (int)((n + COLRED_CH - 1) / COLRED_CH);
    double* partial = workspace(h, BUF_RED, (size_t)2 * nch * m * sizeof(double));
    if (!partial) return -1;
    dim3 grid((unsigned)((m + 127) / 128), (unsigned)nch);
    colred_partial_kernel<<<grid, 128, 0, h->stream>>>(A, lda, (int)n, (int)m, a, partial, partial + (int64_t)nch * m);
    predict_finish_kernel<<<(unsigned)((m + 127) / 128), 128, 0, h->stream>>>(partial, partial + (int64_t)nch * m, nch, (int)m,
                                                                           kdiag, mean, var);
    h->launches += 2;
    return check_cuda(h, cudaGetLastError(), "predict reduce kernels launch");
}

}  // namespace gpb
