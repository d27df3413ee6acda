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
//
// Round 2, in this file as well:
//   factor_inv_chain      the bottom of the recursion (diagonal blocks <= 1024 rows) as a look-ahead chain over
//                         the 128-row leaves: the next leaf waits for one panel product and one 128-column update
//   factor_L, solve_L_*   factor ONLY (N^3/3 flop) for value-only / predict-only flows (GPR/predictor.py:6,
//                         Multi-Input_GPR/main.py:434, BASELINE config C4), block substitution with the inverses of
//                         the 1024-row diagonal blocks only
//   factor_inv_pipelined  right-looking variant over two SM partitions (partition.cu); measured slower, opt-in
#include <stdlib.h>

#include "block_chol.cuh"
#include "engine.cuh"

namespace gpb {

constexpr int NB = 128;
constexpr int NBD = GPB_NBD;   // diagonal-block size of the factor-only path (engine.cuh)
constexpr int LEAF_THREADS = 512;
constexpr size_t LEAF_SMEM = (size_t)(NB * SLD + 64 * TLD + DINV_DOUBLES + 16) * sizeof(double);

// One CTA: L = chol(A_blk) in place, W_blk = L^-1, logdiag[blk] = sum log L_ii, info = first bad pivot.
// The block lives in shared memory (stride SLD); factorisation and inverse run on DMMA tiles
// (block_chol.cuh).  Blocks narrower than 128 are padded to a multiple of 8 with an identity.
__global__ void __launch_bounds__(LEAF_THREADS)
leaf_potrf_inv_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ W, int64_t ldw, int n, int offset,
                      double* __restrict__ logdiag, int* __restrict__ info, int store_L, int info_base) {
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
    block_potrf_inv(S, np, fail, dinv, T);   // L below the diagonal, the tiles of W = L^-1 transposed above it
    if (tid == 0 && *fail != 0) atomicCAS(info, 0, info_base + offset + *fail);
    // L itself is only needed by callers that keep the factor (gpb_potrf, SVGP adjoint); the LML / K^-1
    // pipeline consumes W and the log-diagonal only, so the 128 KB store is skipped there
    if (store_L) {
#pragma unroll 8
        for (int idx = tid; idx < n * NB; idx += LEAF_THREADS) {
            const int i = idx >> 7, j = idx & (NB - 1);
            if (j < n) A[(int64_t)i * lda + j] = (j <= i) ? S[i * SLD + j] : 0.0;
        }
    }
    // sum of the log-diagonal (fixed order, warp 0, from the inverted diagonal blocks) beside the move of W into
    // row-major position by the other warps
    if (store_L) __syncthreads();                 // L read out before W overwrites it
    if (tid < 32) {
        const double s = warp_logdiag_from_dinv(dinv, n);
        if (tid == 0) logdiag[offset / NB] = s;
    }
    block_w_to_lower(S, np, dinv, 1);
#pragma unroll 8
    for (int idx = tid; idx < n * NB; idx += LEAF_THREADS) {
        const int i = idx >> 7, j = idx & (NB - 1);
        if (j < n) W[(int64_t)i * ldw + j] = (j <= i) ? S[i * SLD + j] : 0.0;   // explicit zeros above the diagonal
    }
}

static int leaf(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int n, int offset, double* logdiag,
                int* info, bool store_L, int info_base) {
    static bool attr_set[GPB_MAX_DEVICES] = {};
    {
        cudaError_t e = ensure_dyn_smem(attr_set, h->device, leaf_potrf_inv_kernel, LEAF_SMEM);
        if (e != cudaSuccess) return check_cuda(h, e, "leaf cudaFuncSetAttribute");
    }
    ProfScope prof(h, PROF_LEAF, h->stream);
    if (h->use_pdl) {
        cudaError_t le = launch_pdl(leaf_potrf_inv_kernel, dim3(1), dim3(LEAF_THREADS), LEAF_SMEM, h->stream,
                                    A + (int64_t)offset * lda + offset, lda, W + (int64_t)offset * ldw + offset, ldw, n,
                                    offset, logdiag, info, store_L ? 1 : 0, info_base);
        if (le != cudaSuccess) return check_cuda(h, le, "leaf launch (PDL)");
    } else {
        leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A + (int64_t)offset * lda + offset, lda,
                                                                         W + (int64_t)offset * ldw + offset, ldw, n, offset,
                                                                         logdiag, info, store_L ? 1 : 0, info_base);
    }
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "leaf_potrf_inv_kernel launch");
}

// A, W: full matrices; factor the diagonal block [o, o+n).  keepL: additionally leave L21 in A21
// (needs U scratch of n2 x n1 doubles).
// info_base: added to the reported pivot index (the block may sit at row info_base of a larger matrix whose
// diagonal blocks are factorised one by one, factor_L below).
// ---- the bottom of the recursion as a look-ahead chain -----------------------------------------------------
// A diagonal block of n <= chain_limit rows (1024; 2048 inside matrices of 4096 rows or more): right-looking over
// its 128-row leaves k = 0 .. nb-1,
//   critical stream:  D_k (leaf: L_kk, W_kk)  ->  P_k  T[k+1:, k] = A[k+1:, k] W_kk^T
//                                             ->  Sa_k A[k+1:, k+1] -= T[k+1:, k] T[k+1, k]^T   (next leaf's column)
//   side stream:      Sb_k A[k+2:, k+2:] -= T[k+2:, k] T[k+2:, k]^T ;  R_k  W[k, :k] = -W_kk (T[k, :k] W[:k, :k])
// The 2 x 2 recursion puts EVERY product of a node on the path to the next leaf (its W21 = -W22 U closes the
// node before the parent may use W11): 82 us per leaf at n = 1024 for a 35 us leaf.  Here the next leaf waits
// for one panel product and one 128-column update only; the rest of the trailing update and the rows of the
// inverse run beside the next leaf.  Same flop, same storage (T lives in W's off-diagonal part until R_k
// replaces it by the inverse; V = T W goes through A's dead off-diagonal part, or through scratch when keepL
// keeps L there).  gpb_set_option(h, 5, 0) restores the plain recursion.
static cudaEvent_t chain_event(gpb_handle* h, size_t i) {
    while (h->chain_events.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        h->chain_events.push_back(e);
    }
    return h->chain_events[i];
}

// GPB_CHAIN_MAX (tuning knob): fixed size limit of the blocks the chain takes; < 256 switches it off.  Unset: 1024
// rows, 2048 when the whole matrix has 4096 rows or more (N = 1000 / 2048 / 4096 / 8192 LML+grad: 0.447 / 1.158 /
// 3.851 / 20.62 ms with 1024, 0.449 / 1.225 / 3.757 / 20.38 with 2048, 4.83 / 22.7 at 4096 / 8192 with 4096).
static int chain_env() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("GPB_CHAIN_MAX");
        v = e ? atoi(e) : -1;
        if (v >= 0 && v < 2 * NB) v = 0;
    }
    return v;
}
static int chain_max() {   // upper bound over all matrix sizes (scratch sizing)
    const int e = chain_env();
    return e >= 0 ? e : 2048;
}
static int chain_limit(int64_t total_rows) {
    const int e = chain_env();
    return e >= 0 ? e : (total_rows >= 4096 ? 2048 : 1024);
}

static int factor_inv_chain(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int o, int n, double* logdiag,
                            int* info, bool keepL, double* scratchU, int depth, int info_base) {
    const int nb = (n + NB - 1) / NB;
    const bool fork = h->fork_streams && depth < gpb_handle::MAX_DEPTH && h->side[depth];
    cudaStream_t SA = h->stream, SB = fork ? h->side[depth] : h->stream;
    auto off = [&](int k) { return (int64_t)o + (int64_t)k * NB; };
    auto size = [&](int k) { return (k == nb - 1) ? (n - k * NB) : NB; };
    // events: 2k = P_k done (critical stream), 2k + 1 = Sb_k done (side stream); shared pool, chains never nest
    auto evP = [&](int k) { return chain_event(h, 2 * (size_t)k); };
    auto evSb = [&](int k) { return chain_event(h, 2 * (size_t)k + 1); };
    if (fork && (!evP(nb) || !evSb(nb))) return set_error(h, -1, "chain: event creation failed");
#define GPB_CU(call, what)                                  \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return check_cuda(h, e_, what); \
    } while (0)
    int rc;
    GemmArgs g;
    const int64_t end = (int64_t)o + n;
    for (int k = 0; k < nb; ++k) {
        const int64_t rk = off(k);
        const int nk = size(k);
        double* Wkk = W + rk * ldw + rk;
        if ((rc = leaf(h, A, lda, W, ldw, nk, (int)rk, logdiag, info, keepL, info_base))) return rc;
        if (k + 1 < nb) {
            const int64_t r1 = off(k + 1), m1 = end - r1;
            const int n1 = size(k + 1);
            g = GemmArgs();   // P_k
            g.transa = 0; g.transb = 1; g.M = m1; g.N = nk; g.K = nk;
            g.A = A + r1 * lda + rk; g.lda = lda; g.B = Wkk; g.ldb = ldw; g.C = W + r1 * ldw + rk; g.ldc = ldw; g.b_upper = 1;
            if ((rc = launch_gemm(h, g, SA))) return rc;
            if (fork) GPB_CU(cudaEventRecord(evP(k), SA), "chain record P");
            if (fork && k > 0) GPB_CU(cudaStreamWaitEvent(SA, evSb(k - 1), 0), "chain wait Sb");
            g = GemmArgs();   // Sa_k
            g.transa = 0; g.transb = 1; g.M = m1; g.N = n1; g.K = nk; g.alpha = -1.0; g.beta = 1.0;
            g.A = W + r1 * ldw + rk; g.lda = ldw; g.B = g.A; g.ldb = ldw; g.C = A + r1 * lda + r1; g.ldc = lda;
            if ((rc = launch_gemm(h, g, SA))) return rc;
        }
        if (fork && (k + 1 < nb || k > 0)) {
            if (k + 1 < nb) GPB_CU(cudaStreamWaitEvent(SB, evP(k), 0), "chain wait P");
            else {   // last block: its row of the inverse needs the leaf only
                GPB_CU(cudaEventRecord(evP(k), SA), "chain record D");
                GPB_CU(cudaStreamWaitEvent(SB, evP(k), 0), "chain wait D");
            }
        }
        if (k + 2 < nb) {
            const int64_t r2 = off(k + 2), m2 = end - r2;
            g = GemmArgs();   // Sb_k
            g.transa = 0; g.transb = 1; g.M = m2; g.N = m2; g.K = nk; g.alpha = -1.0; g.beta = 1.0;
            g.A = W + r2 * ldw + rk; g.lda = ldw; g.B = g.A; g.ldb = ldw; g.C = A + r2 * lda + r2; g.ldc = lda; g.tri = 1;
            if ((rc = launch_gemm(h, g, SB))) return rc;
        }
        if (fork && k + 1 < nb) GPB_CU(cudaEventRecord(evSb(k), SB), "chain record Sb");
        if (k > 0) {
            const int64_t w = rk - o;   // width of the finished part of the block
            double* V = keepL ? scratchU : A + rk * lda + o;
            const int64_t ldv = keepL ? w : lda;
            g = GemmArgs();   // V = T[k, :k] W[:k, :k]
            g.transa = 0; g.transb = 0; g.M = nk; g.N = w; g.K = w;
            g.A = W + rk * ldw + o; g.lda = ldw; g.B = W + (int64_t)o * ldw + o; g.ldb = ldw; g.C = V; g.ldc = ldv; g.b_lower = 1;
            if ((rc = launch_gemm(h, g, SB))) return rc;
            if (keepL)   // keep L[k, :k] (= T) under the diagonal of A before the inverse overwrites it in W
                GPB_CU(cudaMemcpy2DAsync(A + rk * lda + o, lda * sizeof(double), W + rk * ldw + o, ldw * sizeof(double),
                                         (size_t)w * sizeof(double), (size_t)nk, cudaMemcpyDeviceToDevice, SB), "chain copy L");
            g = GemmArgs();   // W[k, :k] = -W_kk V
            g.transa = 0; g.transb = 0; g.M = nk; g.N = w; g.K = nk; g.alpha = -1.0;
            g.A = Wkk; g.lda = ldw; g.B = V; g.ldb = ldv; g.C = W + rk * ldw + o; g.ldc = ldw; g.a_lower = 1;
            if ((rc = launch_gemm(h, g, SB))) return rc;
        }
    }
    if (fork && nb > 1) {
        GPB_CU(cudaEventRecord(evSb(nb), SB), "chain record join");
        GPB_CU(cudaStreamWaitEvent(SA, evSb(nb), 0), "chain join");
    }
#undef GPB_CU
    return 0;
}

static int factor_inv_rec(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int o, int n, double* logdiag,
                          int* info, bool keepL, double* scratchU, int depth, int info_base = 0) {
    if (n <= NB) return leaf(h, A, lda, W, ldw, n, o, logdiag, info, keepL, info_base);
    if (h->use_chain && n <= h->chain_limit) return factor_inv_chain(h, A, lda, W, ldw, o, n, logdiag, info, keepL, scratchU, depth, info_base);
    const int n1 = ((n / 2 + NB - 1) / NB) * NB, n2 = n - n1, o2 = o + n1;
    int rc = factor_inv_rec(h, A, lda, W, ldw, o, n1, logdiag, info, keepL, scratchU, depth + 1, info_base);
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
                             depth + 1, info_base)))
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
// scratch doubles keepL needs at size n: the recursion's live U blocks, and -- for a block the look-ahead chain
// may take (the option can change between calls, so both are covered) -- the chain's V block, size(k) x k 128
static size_t keepL_scratch(int n) {
    if (n <= NB) return 0;
    const int n1 = ((n / 2 + NB - 1) / NB) * NB, n2 = n - n1;
    const size_t below = keepL_scratch(n1);
    const size_t right = (size_t)n2 * n1 + keepL_scratch(n2);
    size_t need = below > right ? below : right;
    if (n <= chain_max()) {
        const int nb = (n + NB - 1) / NB;
        for (int k = 1; k < nb; ++k) {
            const size_t v = (size_t)((k == nb - 1) ? (n - k * NB) : NB) * (size_t)(k * NB);
            if (v > need) need = v;
        }
    }
    return need;
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
    h->chain_limit = chain_limit(N);
    return factor_inv_rec(h, A, lda, W, ldw, 0, (int)N, logdiag, d_info, keepL, scratch, 0);
}

// ---- pipelined factor + inverse (+ K^-1) over two SM partitions -------------------------------------------
// The recursion above is latency-bound at its bottom: at N = 8192 the 64 one-CTA leaves and ~240 small
// products are 5.3 of 21.4 ms, most of it exposed, because every product above them needs the finished
// inverse of its whole left neighbour.  A right-looking blocked formulation removes that dependency: with
// block size b (1024) and blocks k = 0 .. nb-1,
//   D_k   L_kk, W_kk = factor_inv(A_kk)                                   the latency-bound chain
//   P_k   T[k+1:, k] = A[k+1:, k] W_kk^T                (= L[k+1:, k], kept in W's storage)
//   Sa_k  A[k+1:, k+1] -= T[k+1:, k] T[k+1, k]^T        next block column only: D_k+1 may start
//   Sb_k  A[k+2:, k+2:] -= T[k+2:, k] T[k+2:, k]^T      rest of the trailing update
//   R_k   V = T[k, :k] W[:k, :k]  (into A[k, :k]);  W[k, :k] = -W_kk V      row k of the inverse
//   Q_k   K^-1[k, :k+1] = W_kk^T W[k, :k+1];  K^-1[:k, :k] += W[k, :k]^T W[k, :k]   (into A, rows <= k are dead)
// D_k+1 depends on Sa_k only, so it runs in the small SM partition (partition.cu) BESIDE Sb_k, R_k, Q_k in the
// bulk partition: two in-order streams, D on one, everything else on the other, two events per step.  Same
// flop count as the recursion (N^3 with K^-1), all of it in products with K = b and large M, N.  D_0 (nothing
// to hide behind) and the last step's R, Q (no D left to run beside them) use the whole device.
// Storage is the recursion's: A ends as the lower tiles of K^-1 (want_kinv) and W as L^-1.
struct StreamSwap {
    gpb_handle* h;
    cudaStream_t s0;
    cudaStream_t side0[gpb_handle::MAX_DEPTH];
    StreamSwap(gpb_handle* h_, cudaStream_t s, const cudaStream_t* sides) : h(h_), s0(h_->stream) {
        h->stream = s;
        for (int i = 0; i < gpb_handle::MAX_DEPTH; ++i) {
            side0[i] = h->side[i];
            h->side[i] = sides[i];
        }
    }
    ~StreamSwap() {
        h->stream = s0;
        for (int i = 0; i < gpb_handle::MAX_DEPTH; ++i) h->side[i] = side0[i];
    }
};

static int pipe_block() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPB_PIPE_B");
        v = e ? atoi(e) : 1024;
        if (v < NB || v % NB != 0) v = 1024;
    }
    return v;
}

bool pipeline_applies(const gpb_handle* h, int64_t N) {
    // (with the stream forks switched off -- bench.py's per-kernel timing -- the same task list runs on the
    // caller's stream alone)
    return h->use_pipeline && (h->part_ok || !h->fork_streams) && N >= 3 * (int64_t)pipe_block();
}

int factor_inv_pipelined(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int64_t N, double* logdiag,
                         int* d_info, bool want_kinv, cudaEvent_t* ev_W_ready) {
    const int b = pipe_block();
    const int nb = (int)((N + b - 1) / b);
    const bool serial = !h->fork_streams || !h->part_ok;
    cudaStream_t S0 = h->stream, M = serial ? h->stream : h->part_bulk, C = serial ? h->stream : h->part_crit;
    auto off = [&](int k) { return (int64_t)k * b; };
    auto size = [&](int k) { return (int)((N - off(k) < b) ? (N - off(k)) : b); };
    auto evD = [&](int k) { return partition_event(h, 2 * (size_t)k); };
    auto evSa = [&](int k) { return partition_event(h, 2 * (size_t)k + 1); };
    cudaEvent_t ev_start = partition_event(h, 2 * (size_t)nb), ev_bulk = partition_event(h, 2 * (size_t)nb + 1),
                ev_W = partition_event(h, 2 * (size_t)nb + 2), ev_crit = partition_event(h, 2 * (size_t)nb + 3);
    if (!ev_start || !ev_bulk || !ev_W || !ev_crit || !evD(nb - 1) || !evSa(nb - 1)) return set_error(h, -1, "pipeline: event creation failed");
    int rc;
    cudaError_t e = cudaMemsetAsync(d_info, 0, sizeof(int), S0);
    if (e != cudaSuccess) return check_cuda(h, e, "memset info");
#define GPB_CU(call, what)                                  \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return check_cuda(h, e_, what); \
    } while (0)
    h->chain_limit = chain_limit(b);
    // D_0 on the caller's stream (whole device)
    if ((rc = factor_inv_rec(h, A, lda, W, ldw, 0, size(0), logdiag, d_info, false, nullptr, 0))) return rc;
    GPB_CU(cudaEventRecord(ev_start, S0), "pipeline start record");
    GPB_CU(cudaStreamWaitEvent(M, ev_start, 0), "pipeline start wait");
    GemmArgs g;
    for (int k = 0; k < nb; ++k) {
        const bool last = (k == nb - 1);
        const int64_t rk = off(k);
        const int nk = size(k);
        cudaStream_t B = last ? S0 : M;
        if (last && nb > 1) GPB_CU(cudaStreamWaitEvent(S0, ev_bulk, 0), "pipeline bulk join");
        if (k > 0) GPB_CU(cudaStreamWaitEvent(B, evD(k), 0), "pipeline wait D");
        double* Wkk = W + rk * ldw + rk;
        if (!last) {
            const int64_t r1 = off(k + 1), m1 = N - r1;
            const int n1 = size(k + 1);
            // P_k: T[k+1:, k] = A[k+1:, k] W_kk^T
            g = GemmArgs();
            g.transa = 0; g.transb = 1; g.M = m1; g.N = nk; g.K = nk;
            g.A = A + r1 * lda + rk; g.lda = lda; g.B = Wkk; g.ldb = ldw; g.C = W + r1 * ldw + rk; g.ldc = ldw; g.b_upper = 1;
            if ((rc = launch_gemm(h, g, M))) return rc;
            // Sa_k: A[k+1:, block k+1] -= T[k+1:, k] T[k+1, k]^T
            g = GemmArgs();
            g.transa = 0; g.transb = 1; g.M = m1; g.N = n1; g.K = nk; g.alpha = -1.0; g.beta = 1.0;
            g.A = W + r1 * ldw + rk; g.lda = ldw; g.B = W + r1 * ldw + rk; g.ldb = ldw; g.C = A + r1 * lda + r1; g.ldc = lda;
            if ((rc = launch_gemm(h, g, M))) return rc;
            GPB_CU(cudaEventRecord(evSa(k), M), "pipeline record Sa");
            // D_k+1 in the small partition
            GPB_CU(cudaStreamWaitEvent(C, evSa(k), 0), "pipeline wait Sa");
            if (serial) {
                rc = factor_inv_rec(h, A, lda, W, ldw, (int)r1, n1, logdiag, d_info, false, nullptr, 0);
            } else {
                StreamSwap swap(h, C, h->part_crit_side);
                rc = factor_inv_rec(h, A, lda, W, ldw, (int)r1, n1, logdiag, d_info, false, nullptr, 0);
            }
            if (rc) return rc;
            GPB_CU(cudaEventRecord(evD(k + 1), C), "pipeline record D");
            // Sb_k: rest of the trailing update
            if (k + 2 < nb) {
                const int64_t r2 = off(k + 2), m2 = N - r2;
                g = GemmArgs();
                g.transa = 0; g.transb = 1; g.M = m2; g.N = m2; g.K = nk; g.alpha = -1.0; g.beta = 1.0;
                g.A = W + r2 * ldw + rk; g.lda = ldw; g.B = g.A; g.ldb = ldw; g.C = A + r2 * lda + r2; g.ldc = lda; g.tri = 1;
                if ((rc = launch_gemm(h, g, M))) return rc;
            }
        }
        if (k > 0) {
            // R_k: V = T[k, :k] W[:k, :k] -> A[k, :k];  W[k, :k] = -W_kk V
            g = GemmArgs();
            g.transa = 0; g.transb = 0; g.M = nk; g.N = rk; g.K = rk;
            g.A = W + rk * ldw; g.lda = ldw; g.B = W; g.ldb = ldw; g.C = A + rk * lda; g.ldc = lda; g.b_lower = 1;
            if ((rc = launch_gemm(h, g, B))) return rc;
            g = GemmArgs();
            g.transa = 0; g.transb = 0; g.M = nk; g.N = rk; g.K = nk; g.alpha = -1.0;
            g.A = Wkk; g.lda = ldw; g.B = A + rk * lda; g.ldb = lda; g.C = W + rk * ldw; g.ldc = ldw; g.a_lower = 1;
            if ((rc = launch_gemm(h, g, B))) return rc;
        }
        if (last) GPB_CU(cudaEventRecord(ev_W, B), "pipeline record W");
        if (want_kinv) {
            if (k > 0) {
                // Q_k: K^-1[k, :k] = W_kk^T W[k, :k] ;  K^-1[:k, :k] += W[k, :k]^T W[k, :k]
                g = GemmArgs();
                g.transa = 1; g.transb = 0; g.M = nk; g.N = rk; g.K = nk;
                g.A = Wkk; g.lda = ldw; g.B = W + rk * ldw; g.ldb = ldw; g.C = A + rk * lda; g.ldc = lda; g.a_upper = 1;
                if ((rc = launch_gemm(h, g, B))) return rc;
                g = GemmArgs();
                g.transa = 1; g.transb = 0; g.M = rk; g.N = rk; g.K = nk; g.beta = 1.0;
                g.A = W + rk * ldw; g.lda = ldw; g.B = g.A; g.ldb = ldw; g.C = A; g.ldc = lda; g.tri = 1;
                if ((rc = launch_gemm(h, g, B))) return rc;
            }
            // K^-1[k, k] = W_kk^T W_kk (lower tiles)
            g = GemmArgs();
            g.transa = 1; g.transb = 0; g.M = nk; g.N = nk; g.K = nk;
            g.A = Wkk; g.lda = ldw; g.B = Wkk; g.ldb = ldw; g.C = A + rk * lda + rk; g.ldc = lda; g.tri = 1; g.a_upper = 1;
            if ((rc = launch_gemm(h, g, B))) return rc;
        }
        if (k == nb - 2) GPB_CU(cudaEventRecord(ev_bulk, M), "pipeline record bulk");
    }
    // the small partition's side streams joined their forks inside the recursion; its main stream ends at D_nb-1,
    // which the last step waited for
    (void)ev_crit;
#undef GPB_CU
    if (ev_W_ready) *ev_W_ready = ev_W;
    return 0;
}

int lauum_lower(gpb_handle* h, const double* d_W, int64_t N, int64_t ldw, double* d_Out, int64_t ldo) {
    GemmArgs g;
    g.transa = 1; g.transb = 0; g.M = N; g.N = N; g.K = N;
    g.A = d_W; g.lda = ldw; g.B = d_W; g.ldb = ldw; g.C = d_Out; g.ldc = ldo; g.tri = 1; g.a_upper = 1;
    return launch_gemm(h, g, h->stream);
}

__global__ void trmv_lower_kernel(const double* __restrict__ W, int64_t ldw, int n, const double* __restrict__ y,
                                  double* __restrict__ out);

// ---- factor only (no N x N inverse) -----------------------------------------------------------------------
// LML-only and predict-only flows (GPR/predictor.py:6, Multi-Input_GPR/main.py:434, BASELINE config C4) need
// L, L^-1 y and L^-1 K(X, X*), not K^-1: N^3/3 flop instead of the 2N^3/3 of factor_inv.  Recursive blocked
// Cholesky whose triangular solves are GEMMs with the explicit inverses of the NBD x NBD DIAGONAL blocks only:
//   potrf(o, n):  n <= NBD: L_kk, Wd_k = factor_inv(A_kk)                                (keepL)
//                 else      potrf(A11); X = A21 L11^-T (trsm, below); A22 -= X X^T; potrf(A22)
//   trsm(B, L):   n <= NBD: X = B Wd^T
//                 else      X1 = trsm(B1, L11); B2 -= X1 L21^T; X2 = trsm(B2, L22)
// Storage: the diagonal blocks of L stay in A; everything below them is written to Lw at the same
// coordinates (a GEMM cannot run in place; A's off-diagonal part is the solve's scratch), so consumers take
// L_kk from A and L_kj (j < k) from Lw.  Wd is an [N, NBD] strip: block k at rows [k NBD, (k+1) NBD).
// Extra flop over N^3/3: N NBD^2 / 3 for the block inverses (1.6 % at N = 8192).
static int trsm_rec(gpb_handle* h, double* A, int64_t lda, double* Lw, int64_t ldl, const double* Wd, int r0, int m, int c0,
                    int n) {
    int rc;
    GemmArgs g;
    if (n <= NBD) {
        g.transa = 0; g.transb = 1; g.M = m; g.N = n; g.K = n;
        g.A = A + (int64_t)r0 * lda + c0; g.lda = lda;
        g.B = Wd + (int64_t)c0 * NBD; g.ldb = NBD;
        g.C = Lw + (int64_t)r0 * ldl + c0; g.ldc = ldl; g.b_upper = 1;
        return launch_gemm(h, g, h->stream);
    }
    const int n1 = ((n / 2 + NBD - 1) / NBD) * NBD, n2 = n - n1;
    if ((rc = trsm_rec(h, A, lda, Lw, ldl, Wd, r0, m, c0, n1))) return rc;
    g.transa = 0; g.transb = 1; g.M = m; g.N = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0;
    g.A = Lw + (int64_t)r0 * ldl + c0; g.lda = ldl;
    g.B = Lw + (int64_t)(c0 + n1) * ldl + c0; g.ldb = ldl;
    g.C = A + (int64_t)r0 * lda + c0 + n1; g.ldc = lda;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    return trsm_rec(h, A, lda, Lw, ldl, Wd, r0, m, c0 + n1, n2);
}

static int potrf_rec(gpb_handle* h, double* A, int64_t lda, double* Lw, int64_t ldl, double* Wd, int o, int n,
                     double* logdiag, int* info, double* scratchU) {
    int rc;
    if (n <= NBD)
        return factor_inv_rec(h, A + (int64_t)o * lda + o, lda, Wd + (int64_t)o * NBD, NBD, 0, n, logdiag + o / NB, info, true,
                              scratchU, 0, o);
    const int n1 = ((n / 2 + NBD - 1) / NBD) * NBD, n2 = n - n1, o2 = o + n1;
    if ((rc = potrf_rec(h, A, lda, Lw, ldl, Wd, o, n1, logdiag, info, scratchU))) return rc;
    if ((rc = trsm_rec(h, A, lda, Lw, ldl, Wd, o2, n2, o, n1))) return rc;
    GemmArgs g;   // A22 -= X X^T (lower tiles)
    g.transa = 0; g.transb = 1; g.M = n2; g.N = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0;
    g.A = Lw + (int64_t)o2 * ldl + o; g.lda = ldl; g.B = g.A; g.ldb = ldl;
    g.C = A + (int64_t)o2 * lda + o2; g.ldc = lda; g.tri = 1;
    if ((rc = launch_gemm(h, g, h->stream))) return rc;
    return potrf_rec(h, A, lda, Lw, ldl, Wd, o2, n2, logdiag, info, scratchU);
}

int factor_L(gpb_handle* h, double* A, int64_t lda, double* Lw, int64_t ldl, double* Wd, int64_t N, double* logdiag,
             int* d_info) {
    double* scratch = nullptr;
    const size_t need = keepL_scratch((int)(N < NBD ? N : NBD));
    if (need) {
        scratch = workspace(h, BUF_PANEL, need * sizeof(double));
        if (!scratch) return -1;
    }
    cudaError_t e = cudaMemsetAsync(d_info, 0, sizeof(int), h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "memset info");
    h->chain_limit = chain_limit(N);
    return potrf_rec(h, A, lda, Lw, ldl, Wd, 0, (int)N, logdiag, d_info, scratch);
}

// out[row] = y[row] - diag * a[row0 + row] - sum_{j < ncols} M[row][j] a[j]   (one warp per row).  Used for the
// off-diagonal strip of a block row (diag = 0) and for the residual y - (K + s2 I) alpha (diag = s2).
__global__ void gemv_sub_kernel(const double* __restrict__ M, int64_t ldm, int nrows, int ncols, const double* __restrict__ a,
                                const double* __restrict__ y, double diag, int row0, double* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const double* l = M + (int64_t)row * ldm;
    double s0 = 0.0, s1 = 0.0;
    const int even = ncols & ~1;
    for (int j = 2 * lane; j < even; j += 64) {     // rows are 16-byte aligned (even leading dimension)
        const double2 lv = *reinterpret_cast<const double2*>(l + j);
        const double2 av = *reinterpret_cast<const double2*>(a + j);
        s0 = fma(lv.x, av.x, s0);
        s1 = fma(lv.y, av.y, s1);
    }
    if (lane == 0 && even < ncols) s0 = fma(l[even], a[even], s0);
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = (y[row] - diag * a[row0 + row]) - s;
}

int gemv_sub(gpb_handle* h, const double* M, int64_t ldm, int64_t nrows, int64_t ncols, const double* a, const double* y,
             double diag, int64_t row0, double* out) {
    const int warps = 8;
    gemv_sub_kernel<<<(unsigned)((nrows + warps - 1) / warps), warps * 32, 0, h->stream>>>(M, ldm, (int)nrows, (int)ncols, a, y,
                                                                                         diag, (int)row0, out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "gemv_sub_kernel launch");
}

// partial[c][j] = sum_{i in row chunk c} M[i][j] v[i]  (thread per column, GEMVT_CH rows per chunk)
constexpr int GEMVT_CH = 32;
__global__ void gemvT_partial_kernel(const double* __restrict__ M, int64_t ldm, int nrows, int ncols,
                                     const double* __restrict__ v, double* __restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= ncols) return;
    const int i0 = c * GEMVT_CH, i1 = min(nrows, i0 + GEMVT_CH);
    double s = 0.0;
    if (i1 - i0 == GEMVT_CH) {
        double w[GEMVT_CH];
#pragma unroll
        for (int k = 0; k < GEMVT_CH; ++k) w[k] = M[(int64_t)(i0 + k) * ldm + j];
#pragma unroll
        for (int k = 0; k < GEMVT_CH; ++k) s = fma(w[k], v[i0 + k], s);
    } else {
        for (int i = i0; i < i1; ++i) s = fma(M[(int64_t)i * ldm + j], v[i], s);
    }
    partial[(int64_t)c * ncols + j] = s;
}
// out[j] = a[j] - sum_c partial[c][j]
__global__ void colsum_sub_kernel(const double* __restrict__ partial, int nchunks, int n, const double* __restrict__ a,
                                  double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(int64_t)c * n + j];
    out[j] = a[j] - s;
}
__global__ void vec_add_kernel(double* __restrict__ x, const double* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += d[i];
}
// out[0] = sum_i x_i y_i  (single block, fixed order)
__global__ void vec_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int n, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s = fma(x[i], y[i], s);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}
int vec_dot(gpb_handle* h, const double* x, const double* y, int64_t n, double* out) {
    vec_dot_kernel<<<1, 256, 0, h->stream>>>(x, y, (int)n, out);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "vec_dot_kernel launch");
}
int vec_add(gpb_handle* h, double* x, const double* d, int64_t n) {
    vec_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(x, d, (int)n);
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "vec_add_kernel launch");
}

// a = L^-1 y by block forward substitution over the NBD blocks of factor_L's output:
// a_k = Wd_k (y_k - sum_{j<k} L_kj a_j).  tmp: N doubles.
int solve_L_vec(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, const double* y, double* a,
                double* tmp) {
    ProfScope prof(h, PROF_VEC, h->stream);
    const int warps = 8;
    for (int64_t o = 0; o < N; o += NBD) {
        const int nk = (int)((N - o < NBD) ? (N - o) : NBD);
        const double* rhs = y + o;
        if (o > 0) {
            gemv_sub_kernel<<<(unsigned)((nk + warps - 1) / warps), warps * 32, 0, h->stream>>>(Lw + o * ldl, ldl, nk, (int)o, a,
                                                                                              y + o, 0.0, 0, tmp + o);
            rhs = tmp + o;
            h->launches += 1;
        }
        trmv_lower_kernel<<<(unsigned)((nk + warps - 1) / warps), warps * 32, 0, h->stream>>>(Wd + o * NBD, NBD, nk, rhs, a + o);
        h->launches += 1;
    }
    return check_cuda(h, cudaGetLastError(), "solve_L_vec launches");
}

// alpha = L^-T a by block back substitution: alpha_k = Wd_k^T (a_k - sum_{i>k} L_ik^T alpha_i).  tmp: N doubles.
int solve_LT_vec(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, const double* a, double* alpha,
                 double* tmp) {
    const int64_t nblocks = (N + NBD - 1) / NBD;
    const int64_t max_chunks = (N + GEMVT_CH - 1) / GEMVT_CH;
    // one request covers the strip partials and trmv_lower_T's own partials (both live in BUF_RED, used in turn)
    double* partial = workspace(h, BUF_RED, (size_t)max_chunks * NBD * sizeof(double));
    if (!partial) return -1;
    int rc;
    for (int64_t k = nblocks - 1; k >= 0; --k) {
        const int64_t o = k * NBD;
        const int nk = (int)((N - o < NBD) ? (N - o) : NBD);
        const int64_t below = N - (o + nk);
        const double* rhs = a + o;
        if (below > 0) {
            ProfScope prof(h, PROF_VEC, h->stream);
            const int nch = (int)((below + GEMVT_CH - 1) / GEMVT_CH);
            dim3 grid((unsigned)((nk + 127) / 128), (unsigned)nch);
            gemvT_partial_kernel<<<grid, 128, 0, h->stream>>>(Lw + (o + nk) * ldl + o, ldl, (int)below, nk, alpha + o + nk, partial);
            colsum_sub_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, h->stream>>>(partial, nch, nk, a + o, tmp + o);
            h->launches += 2;
            rhs = tmp + o;
        }
        if ((rc = trmv_lower_T(h, Wd + o * NBD, NBD, nk, rhs, alpha + o))) return rc;
    }
    return check_cuda(h, cudaGetLastError(), "solve_LT_vec launches");
}

// B [N, m] (ldb) <- destroyed; Out [N, m] (ldo) = L^-1 B by block forward substitution, right-looking:
// Out_k = Wd_k B_k ; B_>k -= L_>k,k Out_k.  N^2 m flop on the DMMA GEMM.  (Right-looking keeps M large: the
// left-looking form updates one 1024-row block per product, 256 tiles on 296 CTA slots -- 27.9 TFLOP/s at
// N = 65536, m = 2048; this form runs the rank-1024 updates over all remaining rows.)
int solve_L_mat(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, double* B, int64_t ldb, int64_t m,
                double* Out, int64_t ldo) {
    int rc;
    for (int64_t o = 0; o < N; o += NBD) {
        const int64_t nk = (N - o < NBD) ? (N - o) : NBD;
        GemmArgs g;
        g.transa = 0; g.transb = 0; g.M = nk; g.N = m; g.K = nk;
        g.A = Wd + o * NBD; g.lda = NBD; g.B = B + o * ldb; g.ldb = ldb; g.C = Out + o * ldo; g.ldc = ldo; g.a_lower = 1;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        const int64_t below = N - (o + nk);
        if (below > 0) {
            g = GemmArgs();
            g.transa = 0; g.transb = 0; g.M = below; g.N = m; g.K = nk; g.alpha = -1.0; g.beta = 1.0;
            g.A = Lw + (o + nk) * ldl + o; g.lda = ldl; g.B = Out + o * ldo; g.ldb = ldo; g.C = B + (o + nk) * ldb; g.ldc = ldb;
            if ((rc = launch_gemm(h, g, h->stream))) return rc;
        }
    }
    return 0;
}

// copy the off-diagonal blocks of L from Lw back under the diagonal blocks in A (gpb_potrf: L in place)
int gather_L(gpb_handle* h, double* A, int64_t lda, const double* Lw, int64_t ldl, int64_t N) {
    for (int64_t o = NBD; o < N; o += NBD) {
        const int64_t nk = (N - o < NBD) ? (N - o) : NBD;
        cudaError_t e = cudaMemcpy2DAsync(A + o * lda, lda * sizeof(double), Lw + o * ldl, ldl * sizeof(double),
                                          (size_t)o * sizeof(double), (size_t)nk, cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) return check_cuda(h, e, "gather_L copy");
    }
    return 0;
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
    if (mean) mean[j] = dt;
    if (var) var[j] = kdiag[j] - ss;
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
