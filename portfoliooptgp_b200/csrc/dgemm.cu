// dgemm.cu -- fp64 GEMM on the FP64 tensor pipe (DMMA.8x8x4, PTX mma.sync.m8n8k4.f64) for sm_100a.
//
// The one genuine dense contraction of the exact-GP path (north_star subsystem 2): every O(N^3)
// step of the Cholesky / inverse / predict pipeline is expressed as calls of this kernel
// (trailing SYRK update, panel solve as a product with the inverted diagonal factor, the TRTRI
// and LAUUM products, W*K(X,X*) for predict_f, and the SVGP M x B products).  It replaces what
// GPflow reaches through tf.linalg.cholesky / triangular_solve / matmul and TF's CholeskyGrad
// (SURVEY.md 2.1 rows K2, K3, K5, K6).
//
// Blackwell note: tcgen05.mma has no f64 kind; on sm_100a every mma.sync f64 shape
// lowers to DMMA.8x8x4 (checked with cuobjdump), measured pipe ceiling 37.2 TFLOP/s, cuBLAS DGEMM
// 35.5 TFLOP/s on this pool (profiles/FP64_PEAKS.json).  Operands are staged with cp.async (LDGSTS)
// into padded shared-memory tiles whose strides make every 64-bit fragment load conflict-free.
//
// Layout: row-major.  C[M,N] = alpha * op(A) op(B) + beta * C.
//   transa = 0: A is [M,K] (k contiguous)      transa = 1: A is [K,M] (m contiguous)
//   transb = 1: B is [N,K] (k contiguous)      transb = 0: B is [K,N] (n contiguous)
// Triangular structure is exploited at tile granularity through per-tile k ranges; diagonal blocks
// of triangular operands must carry explicit zeros in their strict upper part (engine convention).
#include <stdlib.h>

#include "engine.cuh"

namespace gpb {

struct GemmParams {
    const double* A; const double* B; double* C;
    int64_t lda, ldb, ldc;
    int M, N, K;
    double alpha, beta;
    int tri, a_lower, a_upper, b_lower, b_upper;
    int tiles_m, tiles_n;
    int col_major_order, rev_m, rev_n;  // tile enumeration: heaviest k-ranges first (LPT)
    int vecA, vecB, vecC;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tri_index(int b, int& ti, int& tj) {
    int t = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((t + 1) * (t + 2) / 2 <= b) ++t;
    while (t * (t + 1) / 2 > b) --t;
    ti = t;
    tj = b - t * (t + 1) / 2;
}

// Stage one operand tile: R rows (the M or N extent of the tile) by BK k-values.
//   KC = true : global (r, k) at g[r * ld + k]; smem [r][k], stride BK + 4
//   KC = false: global (r, k) at g[k * ld + r]; smem [k][r], stride R + 4
template <int R, int BK, bool KC, int NT>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, int64_t ld, int r0, int rmax, int k0,
                                          int kend, int vec, int tid) {
    if (KC) {
        constexpr int CPR = BK / 2;  // 16-byte chunks per row
        constexpr int LDS = BK + 4;
#pragma unroll
        for (int c = tid; c < R * CPR; c += NT) {
            const int r = c / CPR, kc = (c % CPR) * 2;
            const int gr = r0 + r, gk = k0 + kc;
            double* dst = s + r * LDS + kc;
            const bool v0 = (gr < rmax) && (gk < kend), v1 = (gr < rmax) && (gk + 1 < kend);
            const double* src = g + (int64_t)gr * ld + gk;
            if (vec && v1) {
                cp_async16(dst, src);
            } else {
                cp_async8_zfill(dst, v0 ? src : g, v0);
                cp_async8_zfill(dst + 1, v1 ? src + 1 : g, v1);
            }
        }
    } else {
        constexpr int CPR = R / 2;
        constexpr int LDS = R + 4;
#pragma unroll
        for (int c = tid; c < BK * CPR; c += NT) {
            const int k = c / CPR, rc = (c % CPR) * 2;
            const int gr = r0 + rc, gk = k0 + k;
            double* dst = s + k * LDS + rc;
            const bool v0 = (gr < rmax) && (gk < kend), v1 = (gr + 1 < rmax) && (gk < kend);
            const double* src = g + (int64_t)gk * ld + gr;
            if (vec && v1) {
                cp_async16(dst, src);
            } else {
                cp_async8_zfill(dst, v0 ? src : g, v0);
                cp_async8_zfill(dst + 1, v1 ? src + 1 : g, v1);
            }
        }
    }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKC, bool BKC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
dgemm_kernel(const GemmParams p) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int A_ST = AKC ? BM * (BK + 4) : BK * (BM + 4);
    constexpr int B_ST = BKC ? BN * (BK + 4) : BK * (BN + 4);
    constexpr int LDA_S = AKC ? (BK + 4) : (BM + 4);
    constexpr int LDB_S = BKC ? (BK + 4) : (BN + 4);
    constexpr int MI = WM / 8, NI = WN / 8;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_ST;

    int ti, tj;
    if (p.tri) {
        // lower-triangular tile set with BM = R * BN: row ti holds R (ti + 1) column tiles
        constexpr int R = BM / BN;
        static_assert(BM % BN == 0, "tri mode needs BM a multiple of BN");
        int dummy;
        tri_index((int)(blockIdx.x / R), ti, dummy);
        tj = (int)blockIdx.x - R * (ti * (ti + 1) / 2);
    } else {
        if (p.col_major_order) {
            tj = blockIdx.x / p.tiles_m;
            ti = blockIdx.x % p.tiles_m;
        } else {
            ti = blockIdx.x / p.tiles_n;
            tj = blockIdx.x % p.tiles_n;
        }
        if (p.rev_m) ti = p.tiles_m - 1 - ti;
        if (p.rev_n) tj = p.tiles_n - 1 - tj;
    }
    const int row0 = ti * BM, col0 = tj * BN;
    int kbeg = 0, kend = p.K;
    if (p.a_lower) kend = min(kend, row0 + BM);
    if (p.a_upper) kbeg = max(kbeg, row0);
    if (p.b_lower) kbeg = max(kbeg, col0);
    if (p.b_upper) kend = min(kend, col0 + BN);
    kbeg = (kbeg / BK) * BK;
    const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

    pdl_launch_dependents();   // the next kernel in the stream may start its prologue
    pdl_wait();                // ... and this one waits here for its predecessor's results
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
    const int g = lane >> 2, q = lane & 3;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int kt) {
        if (kt < nk) {
            const int s = kt % STAGES;
            const int k0 = kbeg + kt * BK;
            load_tile<BM, BK, AKC, NT>(As + s * A_ST, p.A, p.lda, row0, p.M, k0, kend, p.vecA, tid);
            load_tile<BN, BK, BKC, NT>(Bs + s * B_ST, p.B, p.ldb, col0, p.N, k0, kend, p.vecB, tid);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);
        const double* as = As + (kt % STAGES) * A_ST;
        const double* bs = Bs + (kt % STAGES) * B_ST;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i)
                a[i] = AKC ? as[(wm0 + i * 8 + g) * LDA_S + kk + q] : as[(kk + q) * LDA_S + wm0 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < NI; ++j)
                b[j] = BKC ? bs[(wn0 + j * 8 + g) * LDB_S + kk + q] : bs[(kk + q) * LDB_S + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row][col..col+1] per 8x8 accumulator tile
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = row0 + wm0 + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = col0 + wn0 + j * 8 + 2 * q;
            if (col >= p.N) continue;
            double* c = p.C + (int64_t)row * p.ldc + col;
            double r0 = alpha * acc[i][j][0], r1 = alpha * acc[i][j][1];
            if (p.vecC && col + 1 < p.N) {
                if (beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2*>(c);
                    r0 = fma(beta, old.x, r0);
                    r1 = fma(beta, old.y, r1);
                }
                *reinterpret_cast<double2*>(c) = make_double2(r0, r1);
            } else {
                if (beta != 0.0) r0 = fma(beta, c[0], r0);
                c[0] = r0;
                if (col + 1 < p.N) {
                    if (beta != 0.0) r1 = fma(beta, c[1], r1);
                    c[1] = r1;
                }
            }
        }
    }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKC, bool BKC>
static int launch_cfg(gpb_handle* h, const GemmParams& p0, cudaStream_t stream) {
    GemmParams p = p0;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int A_ST = AKC ? BM * (BK + 4) : BK * (BM + 4);
    constexpr int B_ST = BKC ? BN * (BK + 4) : BK * (BN + 4);
    constexpr size_t SMEM = (size_t)STAGES * (A_ST + B_ST) * sizeof(double);
    auto kern = dgemm_kernel<BM, BN, BK, WM, WN, STAGES, AKC, BKC>;
    static bool attr_set[GPB_MAX_DEVICES] = {};
    {
        cudaError_t e = ensure_dyn_smem(attr_set, h->device, kern, SMEM);
        if (e != cudaSuccess) return check_cuda(h, e, "dgemm cudaFuncSetAttribute");
    }
    const int tm = (p.M + BM - 1) / BM, tn = (p.N + BN - 1) / BN;
    p.tiles_n = tn;
    p.tiles_m = tm;
    int64_t grid = p.tri ? (int64_t)(BM / BN) * tm * (tm + 1) / 2 : (int64_t)tm * tn;
    if (grid <= 0) return 0;
    const int cat = (BM >= 128) ? PROF_GEMM : PROF_GEMM_SMALL;
    if (h->profile) {
        // flop as executed at tile granularity, to first order: triangular output and triangular operands halve it
        // (lower-triangular output: 1/2; a triangular operand: 1/2; both, as in K^-1 = W^T W: 1/6)
        double f = 2.0 * (double)p.M * (double)p.N * (double)p.K;
        const bool tri_op = p.a_lower || p.a_upper || p.b_lower || p.b_upper;
        f *= (p.tri && tri_op) ? (1.0 / 6.0) : ((p.tri || tri_op) ? 0.5 : 1.0);
        h->prof_flops[cat] += f;
    }
    ProfScope prof(h, cat, stream);
    {
        cudaError_t le = h->use_pdl ? launch_pdl(kern, dim3((unsigned)grid), dim3(NT), SMEM, stream, p)
                                    : (kern<<<(unsigned)grid, NT, SMEM, stream>>>(p), cudaSuccess);
        if (le != cudaSuccess) return check_cuda(h, le, "dgemm_kernel launch (PDL)");
    }
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "dgemm_kernel launch");
}

static int gemm_cfg_override() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPB_GEMM_CFG");   // tuning knob for tools/gemm_bench.py; 0 / unset = automatic
        v = e ? atoi(e) : 0;
    }
    return v;
}

static int big_min_k() {
    static int v = -1;
    if (v < 0) {
        // products with K at or below it never take the 128 x 64 tiles: the rank-128 updates of the look-ahead chain
        // are four k-tiles deep, prologue and epilogue dominate a 128 x 64 CTA (N = 8192 LML+grad: 20.45 ms with
        // them on the large tiles, 20.31 on the 64 x 64 ones; GPB_BIG_MIN_K is the tuning knob)
        const char* e = getenv("GPB_BIG_MIN_K");
        v = e ? atoi(e) : 128;
    }
    return v;
}

static int small_stages() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPB_SMALL_STAGES");   // tuning knob: cp.async stages of the 32 x 32 configuration
        v = e ? atoi(e) : 3;
    }
    return v;
}

template <bool AKC, bool BKC>
static int launch_layout(gpb_handle* h, const GemmParams& p, cudaStream_t stream) {
    switch (gemm_cfg_override()) {
        case 1: return launch_cfg<128, 128, 16, 32, 64, 3, AKC, BKC>(h, p, stream);
        case 2: return launch_cfg<128, 64, 16, 32, 64, 3, AKC, BKC>(h, p, stream);
        case 3: return launch_cfg<128, 128, 32, 32, 64, 2, AKC, BKC>(h, p, stream);
        case 4: return launch_cfg<128, 64, 32, 32, 64, 2, AKC, BKC>(h, p, stream);
        case 5: return launch_cfg<128, 128, 16, 32, 64, 4, AKC, BKC>(h, p, stream);
        default: break;
    }
    // Measured on B200 (tools/gemm_bench.py, 4096^3): 128x64 CTA tiles with two CTAs resident per SM
    // (each hides the other's barrier / prologue / epilogue bubbles) reach 33.5 TFLOP/s = 94% of cuBLAS
    // DGEMM; k-contiguous operands want BK = 32 (256-byte row segments), m/n-contiguous ones BK = 16.
    const int64_t tm128 = (p.M + 127) / 128;
    const int64_t big_tiles = p.tri ? 2 * tm128 * (tm128 + 1) / 2 : tm128 * ((p.N + 63) / 64);
    if (big_tiles >= h->sm_count && p.K > big_min_k()) {
        if (AKC || BKC) return launch_cfg<128, 64, 32, 32, 64, 2, AKC, BKC>(h, p, stream);
        return launch_cfg<128, 64, 16, 32, 64, 3, AKC, BKC>(h, p, stream);
    }
    const int64_t mid_tiles = p.tri ? ((int64_t)((p.M + 63) / 64) * ((p.M + 63) / 64 + 1) / 2)
                                    : ((int64_t)((p.M + 63) / 64) * ((p.N + 63) / 64));
    if (mid_tiles >= h->sm_count)
        return launch_cfg<64, 64, 32, 32, 32, 2, AKC, BKC>(h, p, stream);
    // latency-bound regime (the bottom of the Cholesky recursion): spread over as many SMs as possible
    // and keep the number of dependent load round trips small (BK = 64: K = 128 is two k-tiles; three cp.async
    // stages so that both are in flight from the start -- 2 / 3 / 4 stages: 0.464 / 0.452 / 0.467 ms per LML+grad
    // evaluation at N = 1000, 20.76 / 20.72 / 21.03 at N = 8192)
    switch (small_stages()) {
        case 2: return launch_cfg<32, 32, 64, 16, 16, 2, AKC, BKC>(h, p, stream);
        case 4: return launch_cfg<32, 32, 64, 16, 16, 4, AKC, BKC>(h, p, stream);
        default: return launch_cfg<32, 32, 64, 16, 16, 3, AKC, BKC>(h, p, stream);
    }
}

int launch_gemm(gpb_handle* h, const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return 0;
    if (a.M > 0x7fffffff || a.N > 0x7fffffff || a.K > 0x7fffffff) return set_error(h, -2, "gemm: dimension too large");
    if (a.tri && a.M != a.N) return set_error(h, -2, "gemm: tri needs M == N");
    GemmParams p;
    p.A = a.A; p.B = a.B; p.C = a.C;
    p.lda = a.lda; p.ldb = a.ldb; p.ldc = a.ldc;
    p.M = (int)a.M; p.N = (int)a.N; p.K = (int)a.K;
    p.alpha = a.alpha; p.beta = a.beta;
    p.tri = a.tri; p.a_lower = a.a_lower; p.a_upper = a.a_upper; p.b_lower = a.b_lower; p.b_upper = a.b_upper;
    p.tiles_n = 0; p.tiles_m = 0;
    p.col_major_order = (a.b_upper || a.b_lower) ? 1 : 0;
    p.rev_m = a.a_lower ? 1 : 0;
    p.rev_n = a.b_upper ? 1 : 0;
    auto aligned = [](const void* ptr, int64_t ld) { return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && ((ld & 1) == 0); };
    p.vecA = aligned(a.A, a.lda);
    p.vecB = aligned(a.B, a.ldb);
    p.vecC = aligned(a.C, a.ldc);
    if (a.K <= 0) {
        // C = beta * C: run with an empty k range (nk = 0) -- the epilogue handles it
        p.K = 0;
    }
    const bool akc = (a.transa == 0), bkc = (a.transb != 0);
    if (akc && bkc) return launch_layout<true, true>(h, p, stream);
    if (akc && !bkc) return launch_layout<true, false>(h, p, stream);
    if (!akc && !bkc) return launch_layout<false, false>(h, p, stream);
    return launch_layout<false, true>(h, p, stream);
}

}  // namespace gpb
