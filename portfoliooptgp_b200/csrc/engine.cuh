// engine.cuh -- internal C++ interface of libgpb200: the handle, workspace management and the
// launch wrappers each .cu file provides to capi.cu.  Nothing here crosses the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "kernel_eval.cuh"

struct gpb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    // fork/join streams, one per recursion depth of the blocked factorisation: products that are off
    // the critical path (U = L21 W11) run beside the trailing update and the A22 subtree
    static constexpr int MAX_DEPTH = 10;
    cudaStream_t side[MAX_DEPTH] = {};
    cudaEvent_t ev_fork[MAX_DEPTH] = {}, ev_join[MAX_DEPTH] = {};
    std::string err;
    bool fork_streams = true;   // gpb_set_option(h, 0, x)
    bool use_pdl = true;        // gpb_set_option(h, 1, x): programmatic dependent launch for dgemm / leaf
    bool use_shapes = true;     // gpb_set_option(h, 2, x): straight-line kernels for the known expression shapes (shapes.cuh)
    int refine_mode = 1;        // gpb_set_option(h, 3, x): refinement of y^T K^-1 y in the objective: 0 never, 1 automatic, 2 always (gpr.cu)
    // the factorisation W = L^-1, a = W y left in the workspaces by the last gpr_lml / gpr_predict_f:
    // what it was computed for, and whether the workspaces still hold it (any other use of BUF_K / BUF_W /
    // BUF_VEC clears the flag).  fact_serial counts factorisations; gpb_gpr_predict_f_reuse consumes it.
    bool fact_valid = false;
    bool alpha_valid = false;   // BUF_VEC holds alpha = (K + s2 I)^-1 y of the stored factorisation (gpb_gpr_get_alpha)
    int fact_kind = 0;          // 1: W = L^-1 in BUF_W (factor_inv); 2: factor only -- L in BUF_K (diagonal blocks) / BUF_W, block inverses in BUF_WD
    int64_t fact_serial = 0;
    const double* fact_X = nullptr;
    int64_t fact_N = 0;
    int fact_D = 0;
    double fact_noise = 0.0;
    double fact_theta[GPB_MAX_PARAMS] = {};
    gpb_kernel_spec fact_spec = {};
    int64_t launches = 0;
    int sm_count = 148;
    // SM partitions (green contexts, partition.cu) for the pipelined factorisation: streams of the bulk
    // partition and of the small critical-chain partition; gpb_set_option(h, 4, x) switches the pipeline
    bool part_ok = false;
    bool part_tried = false;
    bool use_chain = true;      // gpb_set_option(h, 5, x): look-ahead chain over the 128-row leaves of a <= chain_limit-row diagonal block
    int chain_limit = 1024;    // rows of the largest block the chain takes in the current factorisation (cholesky.cu)
    bool use_pipeline = false;  // measured slower than the recursion (profiles/r02_pipeline_ab.txt): opt-in
    cudaStream_t part_bulk = nullptr, part_crit = nullptr;
    cudaStream_t part_crit_side[MAX_DEPTH] = {};
    void* part_ctx[2] = {nullptr, nullptr};
    int part_crit_sms = 0, part_bulk_sms = 0;
    std::vector<cudaEvent_t> part_events;
    std::vector<cudaEvent_t> chain_events;   // look-ahead chain at the bottom of the factorisation (cholesky.cu)

    bool has_spec = false;
    gpb_kernel_spec spec;

    // exact-GP binding
    const double* d_X = nullptr;
    const double* d_Yc = nullptr;
    int64_t N = 0;
    int D = 0;

    // grow-only device workspaces
    static constexpr int N_BUF = 10;
    double* buf[N_BUF] = {};
    size_t buf_bytes[N_BUF] = {};
    double* h_pinned = nullptr;  // small pinned staging area for scalar results
    size_t h_pinned_bytes = 0;

    // optional per-category kernel timing with CUDA events on the launch stream (bench.py roofline)
    bool profile = false;
    std::vector<cudaEvent_t> prof_pool;
    struct ProfRec { int cat; int e0, e1; };
    std::vector<ProfRec> prof_recs;
    size_t prof_used = 0;
    double prof_flops[8] = {};   // algorithmic flop of the profiled launches by category (GEMM categories only)
};

namespace gpb {

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may begin (block scheduling,
// prologue) while its predecessor in the stream is still draining; it must call pdl_wait() before it
// touches global memory, and the predecessor calls pdl_launch_dependents() early to allow it.  Used for
// the ~300 short dependent launches at the bottom of the blocked factorisation.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

enum BufId { BUF_K = 0, BUF_W = 1, BUF_VEC = 2, BUF_DINV = 3, BUF_PANEL = 4, BUF_RED = 5, BUF_AUX = 6, BUF_AUX2 = 7, BUF_WD = 8 };
constexpr int GPB_NBD = 2048;   // factor-only path: diagonal blocks of this size carry explicit inverses (cholesky.cu)

// PROF_GEMM: dgemm_kernel launches with the large (128 x 64) tiles -- the throughput-bound products;
// PROF_GEMM_SMALL: its 64 x 64 / 32 x 32 configurations -- the latency-bound bottom of the factorisation
enum ProfCat { PROF_GEMM = 0, PROF_ASSEMBLE = 1, PROF_LEAF = 2, PROF_GRAD = 3, PROF_VEC = 4, PROF_BATCHED = 5, PROF_SVGP = 6, PROF_GEMM_SMALL = 7, PROF_NCAT = 8 };
// RAII timer: records an event pair around the launches issued in its scope when h->profile is on.
struct ProfScope {
    gpb_handle* h; int idx; cudaStream_t st;
    ProfScope(gpb_handle* h_, int cat, cudaStream_t st_);
    ~ProfScope();
};

int set_error(gpb_handle* h, int code, const char* fmt, ...);
int check_cuda(gpb_handle* h, cudaError_t e, const char* what);

// Every C-ABI entry point runs on the handle's device and puts the caller's current device back on
// return (the caller -- torch -- reads it through cudaGetDevice; a process may hold engines on several
// GPUs).  One macro for all translation units: GPB_ENTER(h) declares the guard and returns on failure.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) { prev = -1; return; }
        if (prev != device) err = cudaSetDevice(device);
        else prev = -1;   // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define GPB_ENTER(h)                                                                     \
    if (!(h)) return -1;                                                                 \
    gpb::DeviceGuard gpb_device_guard_((h)->device);                                     \
    if (gpb_device_guard_.err != cudaSuccess) return gpb::check_cuda((h), gpb_device_guard_.err, "cudaSetDevice")
// returns nullptr (and sets error) on failure
double* workspace(gpb_handle* h, int id, size_t bytes);
double* pinned(gpb_handle* h, size_t bytes);

// spec + theta -> device descriptor (host side, validates ranges)
int build_dev_kernel(gpb_handle* h, const double* theta, DevKernel* out);

// ---- assemble.cu
int launch_assemble(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, const double* d_X2, int64_t N2,
                    int D, double* d_K, int64_t ldk, int mode, double diag_add);
int launch_kdiag(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, int D, double* d_out);
// grad reduction: out[p] = sum_{i>=j} c_ij W_ij dK_ij/dtheta_p with W = alpha alpha^T - Kinv (lower stored),
// c_ij = 1 (i == j) or 2 (i > j); also out[n_params] = trace(W) (noise gradient, before the 1/2).
int launch_grad_reduce(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, int D, const double* d_Kinv,
                       int64_t ldk, const double* d_alpha, double* d_out /* n_params + 1, zeroed here */);

// ---- dgemm.cu
// C[M,N] = alpha * op(A) op(B) + beta * C (row-major).  tri: 1 = only tiles with row-block >= col-block.
// k_limit_tri: 0 none; 1 = A is lower-triangular [M,K] (skip k-blocks beyond the row block); see dgemm.cu.
struct GemmArgs {
    int transa = 0, transb = 0;
    int64_t M = 0, N = 0, K = 0;
    double alpha = 1.0, beta = 0.0;
    const double* A = nullptr; int64_t lda = 0;
    const double* B = nullptr; int64_t ldb = 0;
    double* C = nullptr; int64_t ldc = 0;
    int tri = 0;         // 1: compute only lower tiles of C (square M == N)
    int a_lower = 0;     // 1: op(A) [M,K] is lower triangular, K == M: k-range of row block i limited to <= i
    int a_upper = 0;     // 1: op(A) [M,K] is upper triangular, K == M: k-range starts at the row block
    int b_lower = 0;     // 1: op(B) [K,N] is lower triangular, K == N: k-range starts at the col block
    int b_upper = 0;     // 1: op(B) [K,N] is upper triangular: k-range ends at the col block
};
int launch_gemm(gpb_handle* h, const GemmArgs& a, cudaStream_t stream);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per function AND per device: a process that holds
// engines on several GPUs must set it once on each.  One flag per device, per call site (the flag array
// is a function-local static of the templated launcher, i.e. one per kernel instantiation).
constexpr int GPB_MAX_DEVICES = 64;
template <class K>
inline cudaError_t ensure_dyn_smem(bool (&done)[GPB_MAX_DEVICES], int device, K kern, size_t bytes) {
    if (device >= 0 && device < GPB_MAX_DEVICES && done[device]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && device >= 0 && device < GPB_MAX_DEVICES) done[device] = true;
    return e;
}

// ---- partition.cu
void partitions_create(gpb_handle* h);
void partitions_destroy(gpb_handle* h);
cudaEvent_t partition_event(gpb_handle* h, size_t i);

// ---- cholesky.cu
// A (lower, in place) -> L on the diagonal blocks (and everywhere when keepL), W = L^-1 (lower, zero
// strict-upper inside 128-aligned diagonal blocks), logdiag[b] = sum of log L_ii over block b,
// *d_info = 1-based index of the first non-positive pivot (0 = ok).
int factor_inv(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int64_t N, double* logdiag, int* d_info,
               bool keepL);
// Pipelined variant over the two SM partitions (see cholesky.cu); A ends as K^-1 (lower tiles) when want_kinv.
// *ev_W_ready (optional) is recorded on the caller's stream when W is complete (before the last K^-1 products).
bool pipeline_applies(const gpb_handle* h, int64_t N);
int factor_inv_pipelined(gpb_handle* h, double* A, int64_t lda, double* W, int64_t ldw, int64_t N, double* logdiag,
                         int* d_info, bool want_kinv, cudaEvent_t* ev_W_ready);
// Factor only (N^3/3): L's diagonal NBD-blocks in A, the rest in Lw, block inverses in the [N, NBD] strip Wd.
int factor_L(gpb_handle* h, double* A, int64_t lda, double* Lw, int64_t ldl, double* Wd, int64_t N, double* logdiag,
             int* d_info);
int solve_L_vec(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, const double* y, double* a,
                double* tmp);
int solve_LT_vec(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, const double* a, double* alpha,
                 double* tmp);
int gemv_sub(gpb_handle* h, const double* M, int64_t ldm, int64_t nrows, int64_t ncols, const double* a, const double* y,
             double diag, int64_t row0, double* out);
int vec_add(gpb_handle* h, double* x, const double* d, int64_t n);
int vec_dot(gpb_handle* h, const double* x, const double* y, int64_t n, double* out);
int solve_L_mat(gpb_handle* h, const double* Lw, int64_t ldl, const double* Wd, int64_t N, double* B, int64_t ldb, int64_t m,
                double* Out, int64_t ldo);
int gather_L(gpb_handle* h, double* A, int64_t lda, const double* Lw, int64_t ldl, int64_t N);
// Out (lower tiles) = W^T W
int lauum_lower(gpb_handle* h, const double* d_W, int64_t N, int64_t ldw, double* d_Out, int64_t ldo);
int trmv_lower(gpb_handle* h, const double* W, int64_t ldw, int64_t n, const double* y, double* out);
int trmv_lower_T(gpb_handle* h, const double* W, int64_t ldw, int64_t n, const double* a, double* out);
int quad_logdet(gpb_handle* h, const double* v, int64_t n, const double* logdiag, double* out2);
// var[j] = kdiag[j] - sum_i A[i][j]^2 ; dotp[j] = sum_i A[i][j] a[i].  Either output may be null.
int predict_colreduce(gpb_handle* h, const double* A, int64_t lda, int64_t n, int64_t m, const double* a,
                      const double* kdiag, double* dotp, double* var);

// ---- gpr.cu
int gpr_lml(gpb_handle* h, const double* theta, double noise, double* lml, double* grad_theta, double* grad_noise,
            int want_grad);
int gpr_predict_f(gpb_handle* h, const double* theta, double noise, const double* d_Xs, int64_t Ns, double* d_mean,
                  double* d_var, int64_t reuse_serial = -1);

// ---- batched.cu: one GP per CTA.  mode 0 = LML, 1 = LML + gradient, 2 = predict_f
// d_nrows (optional, [B] ints): GP b uses the first d_nrows[b] of its N rows (ragged batches)
int launch_batched(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta, const double* d_noise,
                   const int* d_nrows, int64_t B, int64_t N, int D, int mode, double* d_out, int* d_info, const double* d_Xs,
                   int64_t Ns, double* d_mean, double* d_var);

// ---- svgp.cu
int svgp_data_term(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D,
                   const double* d_qmu, const double* d_Lq, int64_t ldq, const double* d_X, const double* d_y, int64_t B,
                   double* d_flat, int want_grad);
int svgp_finish(gpb_handle* h, double* d_flat, double scale, const double* d_qmu, const double* d_Lq, int64_t ldq,
                int64_t M, int D, int P, int apply_grad, double* h_elbo, double* h_kl);
int svgp_predict_f(gpb_handle* h, const double* theta, const double* d_Z, int64_t M, int D, const double* d_qmu,
                   const double* d_Lq, int64_t ldq, const double* d_Xs, int64_t Ns, double* d_mean, double* d_var);
int sgpr_elbo(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D, const double* d_X,
              const double* d_err, int64_t N, int want_grad, double* h_out, double* d_errbar);
int sgpr_predict_f(gpb_handle* h, const double* theta, double s2, const double* d_Z, int64_t M, int D, const double* d_X,
                   const double* d_err, int64_t N, const double* d_Xs, int64_t Ns, double* d_mean, double* d_var);

}  // namespace gpb
