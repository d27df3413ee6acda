// capi.cu -- the extern "C" surface of libgpb200.so (include/gpb200.h): handle lifecycle, error
// reporting, workspace management, kernel-spec validation, and thin forwarding to the engine.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <thread>
#include <vector>

#include "engine.cuh"
#include "shapes.cuh"

namespace gpb {

int set_error(gpb_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

int check_cuda(gpb_handle* h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return set_error(h, -100 - (int)e, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}

double* workspace(gpb_handle* h, int id, size_t bytes) {
    // whoever asks for these three is about to overwrite the stored GPR factorisation (gpr.cu re-validates
    // it after its own requests)
    if (id == BUF_K || id == BUF_W || id == BUF_VEC || id == BUF_WD) h->fact_valid = false;
    if (bytes <= h->buf_bytes[id] && h->buf[id]) return h->buf[id];
    if (h->buf[id]) {
        cudaFree(h->buf[id]);  // synchronises the device: no kernel can still be using the old block
        h->buf[id] = nullptr;
        h->buf_bytes[id] = 0;
    }
    // grow with head-room so that rolling-window refits (N, N+1, ...) do not reallocate every call
    size_t want = bytes + bytes / 8 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        check_cuda(h, e, "workspace cudaMalloc");
        return nullptr;
    }
    h->buf[id] = static_cast<double*>(p);
    h->buf_bytes[id] = want;
    return h->buf[id];
}

double* pinned(gpb_handle* h, size_t bytes) {
    if (bytes <= h->h_pinned_bytes && h->h_pinned) return h->h_pinned;
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned = nullptr;
    h->h_pinned_bytes = 0;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
        check_cuda(h, e, "cudaMallocHost");
        return nullptr;
    }
    h->h_pinned = static_cast<double*>(p);
    h->h_pinned_bytes = bytes;
    return h->h_pinned;
}

ProfScope::ProfScope(gpb_handle* h_, int cat, cudaStream_t st_) : h(h_), idx(-1), st(st_) {
    if (!h->profile) return;
    while (h->prof_pool.size() < h->prof_used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->prof_pool.push_back(e);
    }
    gpb_handle::ProfRec r;
    r.cat = cat; r.e0 = (int)h->prof_used; r.e1 = (int)h->prof_used + 1;
    h->prof_used += 2;
    cudaEventRecord(h->prof_pool[r.e0], st);
    h->prof_recs.push_back(r);
    idx = (int)h->prof_recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof_pool[h->prof_recs[idx].e1], st);
}

static int validate_spec(gpb_handle* h, const gpb_kernel_spec* s) {
    if (s->n_dims < 1 || s->n_dims > GPB_MAX_DIMS) return set_error(h, -2, "kernel spec: n_dims=%d outside [1,%d]", s->n_dims, GPB_MAX_DIMS);
    if (s->n_params < 1 || s->n_params > GPB_MAX_PARAMS) return set_error(h, -2, "kernel spec: n_params=%d outside [1,%d]", s->n_params, GPB_MAX_PARAMS);
    if (s->n_groups < 1 || s->n_groups > GPB_MAX_GROUPS) return set_error(h, -2, "kernel spec: n_groups=%d outside [1,%d]", s->n_groups, GPB_MAX_GROUPS);
    if (s->n_leaves < 1 || s->n_leaves > GPB_MAX_LEAVES) return set_error(h, -2, "kernel spec: n_leaves=%d outside [1,%d]", s->n_leaves, GPB_MAX_LEAVES);
    if (s->n_terms < 1 || s->n_terms > GPB_MAX_TERMS) return set_error(h, -2, "kernel spec: n_terms=%d outside [1,%d]", s->n_terms, GPB_MAX_TERMS);
    auto idx_ok = [&](int i) { return i >= 0 && i < s->n_params; };
    for (int g = 0; g < s->n_groups; ++g) {
        const gpb_group& G = s->groups[g];
        if (G.kind < 0 || G.kind > GPB_GROUP_DOT) return set_error(h, -2, "kernel spec: group %d bad kind %d", g, G.kind);
        if (G.dim_mask == 0 || (s->n_dims < 32 && (G.dim_mask >> s->n_dims) != 0))
            return set_error(h, -2, "kernel spec: group %d dim_mask 0x%x invalid for D=%d", g, G.dim_mask, s->n_dims);
        const int nact = __builtin_popcount(G.dim_mask);
        if (G.ard_index >= 0 && (!idx_ok(G.ard_index) || !idx_ok(G.ard_index + nact - 1)))
            return set_error(h, -2, "kernel spec: group %d ARD index out of range", g);
        const bool per = (G.kind == GPB_GROUP_PERIODIC_SQ || G.kind == GPB_GROUP_PERIODIC_ABS);
        if (per != (G.period_index >= 0)) return set_error(h, -2, "kernel spec: group %d period index inconsistent", g);
        if (per && !idx_ok(G.period_index)) return set_error(h, -2, "kernel spec: group %d period index out of range", g);
    }
    for (int l = 0; l < s->n_leaves; ++l) {
        const gpb_leaf& L = s->leaves[l];
        if (L.kind < 0 || L.kind > GPB_LEAF_LINEAR) return set_error(h, -2, "kernel spec: leaf %d bad kind %d", l, L.kind);
        if (L.group < 0 || L.group >= s->n_groups) return set_error(h, -2, "kernel spec: leaf %d bad group", l);
        if (!idx_ok(L.var_index)) return set_error(h, -2, "kernel spec: leaf %d variance index out of range", l);
        const gpb_group& G = s->groups[L.group];
        if (L.kind == GPB_LEAF_LINEAR) {
            if (G.kind != GPB_GROUP_DOT || L.ls_index >= 0) return set_error(h, -2, "kernel spec: Linear leaf %d needs a DOT group and no lengthscale", l);
        } else {
            if (G.kind == GPB_GROUP_DOT) return set_error(h, -2, "kernel spec: stationary leaf %d on a DOT group", l);
            if ((L.ls_index >= 0) == (G.ard_index >= 0)) return set_error(h, -2, "kernel spec: leaf %d needs exactly one of scalar/ARD lengthscale", l);
            if (L.ls_index >= 0 && !idx_ok(L.ls_index)) return set_error(h, -2, "kernel spec: leaf %d lengthscale index out of range", l);
            const bool r_kind = (L.kind >= GPB_LEAF_MATERN12 && L.kind <= GPB_LEAF_MATERN52);
            if (G.kind == GPB_GROUP_PERIODIC_ABS && !r_kind) return set_error(h, -2, "kernel spec: leaf %d: PERIODIC_ABS feeds K_r kernels only", l);
            if (G.kind == GPB_GROUP_PERIODIC_SQ && r_kind) return set_error(h, -2, "kernel spec: leaf %d: K_r kernels need PERIODIC_ABS", l);
        }
        if ((L.kind == GPB_LEAF_RQ) != (L.alpha_index >= 0)) return set_error(h, -2, "kernel spec: leaf %d alpha index inconsistent", l);
        if (L.alpha_index >= 0 && !idx_ok(L.alpha_index)) return set_error(h, -2, "kernel spec: leaf %d alpha index out of range", l);
    }
    for (int t = 0; t < s->n_terms; ++t) {
        const gpb_term& T = s->terms[t];
        if (T.n_factors < 1 || T.n_factors > GPB_MAX_FACTORS) return set_error(h, -2, "kernel spec: term %d has %d factors", t, T.n_factors);
        for (int f = 0; f < T.n_factors; ++f)
            if (T.leaf[f] < 0 || T.leaf[f] >= s->n_leaves) return set_error(h, -2, "kernel spec: term %d factor %d bad leaf", t, f);
    }
    return 0;
}

int build_dev_kernel(gpb_handle* h, const double* theta, DevKernel* out) {
    if (!h->has_spec) return set_error(h, -3, "no kernel set (gpb_set_kernel)");
    const gpb_kernel_spec& s = h->spec;
    for (int p = 0; p < s.n_params; ++p)
        if (!(theta[p] == theta[p])) return set_error(h, -2, "theta[%d] is NaN", p);
    memset(out, 0, sizeof(DevKernel));
    const int bad = build_dev_kernel_core(s, theta, out);
    if (bad) return set_error(h, -2, "lengthscale theta[%d]=%g must be > 0", bad - 1, theta[bad - 1]);
    return 0;
}

}  // namespace gpb

using namespace gpb;

extern "C" {

int gpb_version(void) { return 100; }

int gpb_create(gpb_handle** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return -10;  // no CUDA device: there is no CPU fallback
    if (device < 0 || device >= count) return -2;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return -11;
    gpb_handle* h = new (std::nothrow) gpb_handle();
    if (!h) return -12;
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        h->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) {
            delete h;
            return -13;  // built for sm_100a only
        }
    }
    for (int d = 0; d < gpb_handle::MAX_DEPTH; ++d) {
        cudaStreamCreateWithFlags(&h->side[d], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&h->ev_fork[d], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_join[d], cudaEventDisableTiming);
    }
    // (the SM partitions of the opt-in pipelined factorisation are created when option 4 is switched on)
    *out = h;
    return 0;
}

int gpb_destroy(gpb_handle* h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    partitions_destroy(h);
    for (cudaEvent_t e : h->chain_events) cudaEventDestroy(e);
    for (int i = 0; i < gpb_handle::N_BUF; ++i)
        if (h->buf[i]) cudaFree(h->buf[i]);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
    for (int d = 0; d < gpb_handle::MAX_DEPTH; ++d) {
        if (h->side[d]) cudaStreamDestroy(h->side[d]);
        if (h->ev_fork[d]) cudaEventDestroy(h->ev_fork[d]);
        if (h->ev_join[d]) cudaEventDestroy(h->ev_join[d]);
    }
    delete h;
    return 0;
}

const char* gpb_last_error(gpb_handle* h) { return h ? h->err.c_str() : "null handle"; }

int gpb_set_stream(gpb_handle* h, void* cuda_stream) {
    if (!h) return -1;
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int64_t gpb_launch_count(gpb_handle* h) { return h ? h->launches : -1; }

int gpb_kernel_shape(gpb_handle* h) {
    if (!h) return -1;
    if (!h->has_spec) return set_error(h, -3, "kernel_shape: no kernel set (gpb_set_kernel)");
    if (!h->use_shapes) return 0;
    double ones[GPB_MAX_PARAMS];
    for (int p = 0; p < GPB_MAX_PARAMS; ++p) ones[p] = 1.0;
    DevKernel probe;
    build_dev_kernel_core(h->spec, ones, &probe);
    return match_shape(probe);
}

int gpb_set_option(gpb_handle* h, int option, int value) {
    if (!h) return -1;
    switch (option) {
        case 0: h->fork_streams = (value != 0); return 0;
        case 1: h->use_pdl = (value != 0); return 0;
        case 2: h->use_shapes = (value != 0); return 0;
        case 4: {
            h->use_pipeline = (value != 0);
            if (h->use_pipeline && !h->part_ok && !h->part_tried) {
                DeviceGuard guard(h->device);
                h->part_tried = true;
                partitions_create(h);
            }
            return 0;
        }
        case 5: h->use_chain = (value != 0); return 0;
        case 3:
            if (value < 0 || value > 2) return set_error(h, -2, "option 3 (objective refinement) takes 0, 1 or 2");
            h->refine_mode = value;
            return 0;
        default: return set_error(h, -2, "unknown option %d", option);
    }
}

int gpb_profile_enable(gpb_handle* h, int on) {
    if (!h) return -1;
    h->profile = (on != 0);
    h->prof_recs.clear();
    h->prof_used = 0;
    for (int c = 0; c < PROF_NCAT; ++c) h->prof_flops[c] = 0.0;
    return 0;
}

int gpb_profile_read_flops(gpb_handle* h, double* h_flops) {
    if (!h || !h_flops) return -1;
    for (int c = 0; c < PROF_NCAT; ++c) {
        h_flops[c] = h->prof_flops[c];
        h->prof_flops[c] = 0.0;
    }
    return 0;
}

int gpb_profile_read(gpb_handle* h, double* h_ms, int64_t* h_counts) {
    GPB_ENTER(h);
    if (!h_ms || !h_counts) return set_error(h, -2, "profile_read: null pointer");
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "profile_read sync");
    for (int c = 0; c < PROF_NCAT; ++c) { h_ms[c] = 0.0; h_counts[c] = 0; }
    for (const auto& r : h->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->prof_pool[r.e0], h->prof_pool[r.e1]) == cudaSuccess) {
            h_ms[r.cat] += ms;
            h_counts[r.cat] += 1;
        }
    }
    h->prof_recs.clear();
    h->prof_used = 0;
    return 0;
}

int gpb_set_kernel(gpb_handle* h, const gpb_kernel_spec* spec) {
    if (!h || !spec) return -1;
    int rc = validate_spec(h, spec);
    if (rc) return rc;
    h->spec = *spec;
    h->has_spec = true;
    return 0;
}

int gpb_assemble(gpb_handle* h, const double* h_theta, const double* d_X, int64_t N, const double* d_X2, int64_t N2,
                 int D, double* d_K, int64_t ldk, int mode, double diag_add) {
    GPB_ENTER(h);
    if (!h_theta || !d_X || !d_K) return set_error(h, -2, "assemble: null pointer");
    if (mode < 0 || mode > 2) return set_error(h, -2, "assemble: bad mode %d", mode);
    if (!d_X2) { d_X2 = d_X; N2 = N; }
    if (ldk < N2) return set_error(h, -2, "assemble: ldk < N2");
    DevKernel kp;
    int rc = build_dev_kernel(h, h_theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "assemble: kernel expects D=%d, got %d", kp.n_dims, D);
    return launch_assemble(h, kp, d_X, N, d_X2, N2, D, d_K, ldk, mode, diag_add);
}

int gpb_kdiag(gpb_handle* h, const double* h_theta, const double* d_X, int64_t N, int D, double* d_out) {
    GPB_ENTER(h);
    if (!h_theta || !d_X || !d_out) return set_error(h, -2, "kdiag: null pointer");
    DevKernel kp;
    int rc = build_dev_kernel(h, h_theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != D) return set_error(h, -2, "kdiag: kernel expects D=%d, got %d", kp.n_dims, D);
    return launch_kdiag(h, kp, d_X, N, D, d_out);
}

static int potrf_common(gpb_handle* h, double* d_A, int64_t N, int64_t lda, double* d_W, int64_t ldw) {
    if (!d_A || N <= 0 || lda < N) return set_error(h, -2, "potrf: bad arguments");
    if (N > 0x7fffffff) return set_error(h, -2, "potrf: N too large");
    const int64_t nblk = (N + 127) / 128;
    double* v = workspace(h, BUF_DINV, (size_t)(nblk + 16) * sizeof(double));
    if (!v) return -1;
    int* info = reinterpret_cast<int*>(v + nblk);
    int rc = factor_inv(h, d_A, lda, d_W, ldw, N, v, info, true);
    if (rc) return rc;
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp + 64, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "potrf sync");
    const int inf = *reinterpret_cast<int*>(hp + 64);
    if (inf > 0) set_error(h, inf, "Cholesky decomposition was not successful: non-positive pivot at row %d", inf);
    return inf;
}

int gpb_potrf(gpb_handle* h, double* d_A, int64_t N, int64_t lda) {
    GPB_ENTER(h);
    // the factor alone (N^3/3 flop): no N x N inverse; the solves inside multiply by the inverses of the
    // GPB_NBD x GPB_NBD diagonal blocks only (cholesky.cu factor_L)
    if (!d_A || N <= 0 || lda < N) return set_error(h, -2, "potrf: bad arguments");
    if (N > 0x7fffff00LL) return set_error(h, -2, "potrf: N too large");
    const int64_t ldw = (N + 15) / 16 * 16, rows = (N + 127) / 128 * 128, nblk = (N + 127) / 128;
    double* Lw = workspace(h, BUF_W, (size_t)rows * ldw * sizeof(double));
    double* Wd = workspace(h, BUF_WD, (size_t)rows * GPB_NBD * sizeof(double));
    double* v = workspace(h, BUF_DINV, (size_t)(nblk + 16) * sizeof(double));
    if (!Lw || !Wd || !v) return -1;
    int* info = reinterpret_cast<int*>(v + nblk);
    int rc = factor_L(h, d_A, lda, Lw, ldw, Wd, N, v, info);
    if (rc) return rc;
    if ((rc = gather_L(h, d_A, lda, Lw, ldw, N))) return rc;
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp + 64, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "potrf sync");
    const int inf = *reinterpret_cast<int*>(hp + 64);
    if (inf > 0) set_error(h, inf, "Cholesky decomposition was not successful: non-positive pivot at row %d", inf);
    return inf;
}

int gpb_potrf_inv(gpb_handle* h, double* d_A, int64_t N, int64_t lda, double* d_W, int64_t ldw) {
    GPB_ENTER(h);
    if (!d_W || ldw < N) return set_error(h, -2, "potrf_inv: bad W");
    return potrf_common(h, d_A, N, lda, d_W, ldw);
}

int gpb_lauum(gpb_handle* h, const double* d_W, int64_t N, int64_t ldw, double* d_Out, int64_t ldo) {
    GPB_ENTER(h);
    if (!d_W || !d_Out || N <= 0 || ldw < N || ldo < N) return set_error(h, -2, "lauum: bad arguments");
    return lauum_lower(h, d_W, N, ldw, d_Out, ldo);
}

int gpb_gemm(gpb_handle* h, int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha, const double* d_A,
             int64_t lda, const double* d_B, int64_t ldb, double beta, double* d_C, int64_t ldc, int tri) {
    GPB_ENTER(h);
    if (!d_A || !d_B || !d_C) return set_error(h, -2, "gemm: null pointer");
    GemmArgs g;
    g.transa = transa; g.transb = transb; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = d_A; g.lda = lda; g.B = d_B; g.ldb = ldb; g.C = d_C; g.ldc = ldc; g.tri = tri;
    return launch_gemm(h, g, h->stream);
}

int gpb_gpr_set_data(gpb_handle* h, const double* d_X, int64_t N, int D, const double* d_Yc) {
    if (!h) return -1;
    if (!d_X || !d_Yc || N <= 0) return set_error(h, -2, "set_data: bad arguments");
    if (D < 1 || D > GPB_MAX_DIMS) return set_error(h, -2, "set_data: D=%d outside [1,%d]", D, GPB_MAX_DIMS);
    if (N > 0x7fffff00LL) return set_error(h, -2, "set_data: N too large");
    h->d_X = d_X; h->d_Yc = d_Yc; h->N = N; h->D = D;
    return 0;
}

int gpb_gpr_lml(gpb_handle* h, const double* h_theta, double noise_variance, double* h_lml) {
    GPB_ENTER(h);
    if (!h_theta || !h_lml) return set_error(h, -2, "gpr_lml: null pointer");
    return gpr_lml(h, h_theta, noise_variance, h_lml, nullptr, nullptr, 0);
}

int gpb_gpr_lml_grad(gpb_handle* h, const double* h_theta, double noise_variance, double* h_lml, double* h_grad_theta,
                     double* h_grad_noise) {
    GPB_ENTER(h);
    if (!h_theta || !h_lml || !h_grad_theta || !h_grad_noise) return set_error(h, -2, "gpr_lml_grad: null pointer");
    return gpr_lml(h, h_theta, noise_variance, h_lml, h_grad_theta, h_grad_noise, 1);
}

int gpb_gpr_predict_f(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Xs, int64_t Ns,
                      double* d_mean, double* d_var) {
    GPB_ENTER(h);
    if (!h_theta || !d_Xs || !d_mean || !d_var) return set_error(h, -2, "predict_f: null pointer");
    return gpr_predict_f(h, h_theta, noise_variance, d_Xs, Ns, d_mean, d_var);
}

int gpb_gpr_lml_grad_many(int nh, gpb_handle* const* handles, int64_t njobs, const double* const* d_X, const int64_t* N,
                          int D, const double* const* d_Yc, const double* h_theta, int P, const double* h_noise,
                          int want_grad, double* h_lml, double* h_grad_theta, double* h_grad_noise, int* h_rc) {
    if (nh < 1 || !handles || njobs < 0 || !d_X || !N || !d_Yc || !h_theta || !h_noise || !h_lml || !h_rc || P < 0) return -2;
    if (want_grad && (!h_grad_theta || !h_grad_noise)) return -2;
    for (int t = 0; t < nh; ++t) {
        if (!handles[t]) return -2;
        for (int u = 0; u < t; ++u)
            if (handles[u] == handles[t]) return -2;      // one host thread per handle
    }
    auto run = [&](int t) {
        gpb_handle* h = handles[t];
        for (int64_t j = t; j < njobs; j += nh) {
            int rc = gpb_gpr_set_data(h, d_X[j], N[j], D, d_Yc[j]);
            if (rc == 0) {
                rc = want_grad ? gpb_gpr_lml_grad(h, h_theta + j * P, h_noise[j], h_lml + j, h_grad_theta + j * P, h_grad_noise + j)
                               : gpb_gpr_lml(h, h_theta + j * P, h_noise[j], h_lml + j);
            }
            h_rc[j] = rc;
        }
    };
    const int nt = (int)(njobs < nh ? njobs : nh);
    std::vector<std::thread> pool;
    pool.reserve(nt > 0 ? nt - 1 : 0);
    try {
        for (int t = 1; t < nt; ++t) pool.emplace_back(run, t);
    } catch (...) {
        for (auto& th : pool) th.join();
        return -1;
    }
    if (nt > 0) run(0);                                   // the caller's thread takes handle 0
    for (auto& th : pool) th.join();
    return 0;
}

int gpb_gpr_predict_f_many(int nh, gpb_handle* const* handles, int64_t njobs, const double* const* d_X, const int64_t* N,
                           int D, const double* const* d_Yc, const double* h_theta, int P, const double* h_noise,
                           const double* const* d_Xs, int64_t Ns, double* const* d_mean, double* const* d_var, int* h_rc) {
    if (nh < 1 || !handles || njobs < 0 || !d_X || !N || !d_Yc || !h_theta || !h_noise || !d_Xs || !d_mean || !d_var || !h_rc ||
        P < 0 || Ns < 1)
        return -2;
    for (int t = 0; t < nh; ++t) {
        if (!handles[t]) return -2;
        for (int u = 0; u < t; ++u)
            if (handles[u] == handles[t]) return -2;
    }
    auto run = [&](int t) {
        gpb_handle* h = handles[t];
        for (int64_t j = t; j < njobs; j += nh) {
            int rc = gpb_gpr_set_data(h, d_X[j], N[j], D, d_Yc[j]);
            if (rc == 0) rc = gpb_gpr_predict_f(h, h_theta + j * P, h_noise[j], d_Xs[j], Ns, d_mean[j], d_var[j]);
            h_rc[j] = rc;
        }
        DeviceGuard guard(h->device);
        cudaStreamSynchronize(h->stream);
    };
    const int nt = (int)(njobs < nh ? njobs : nh);
    std::vector<std::thread> pool;
    pool.reserve(nt > 0 ? nt - 1 : 0);
    try {
        for (int t = 1; t < nt; ++t) pool.emplace_back(run, t);
    } catch (...) {
        for (auto& th : pool) th.join();
        return -1;
    }
    if (nt > 0) run(0);
    for (auto& th : pool) th.join();
    return 0;
}

int64_t gpb_gpr_factor_serial(gpb_handle* h) { return h ? h->fact_serial : -1; }

int gpb_gpr_predict_f_reuse(gpb_handle* h, const double* h_theta, double noise_variance, int64_t factor_serial,
                            const double* d_Xs, int64_t Ns, double* d_mean, double* d_var) {
    GPB_ENTER(h);
    if (!h_theta || !d_Xs || !d_mean || !d_var) return set_error(h, -2, "predict_f_reuse: null pointer");
    return gpr_predict_f(h, h_theta, noise_variance, d_Xs, Ns, d_mean, d_var, factor_serial);
}

}  // extern "C"

extern "C" {

int gpb_batched_lml_grad(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                         const double* d_noise, int64_t B, int64_t N, int D, double* d_out, int32_t* d_info,
                         int want_grad) {
    GPB_ENTER(h);
    if (!d_X || !d_Yc || !d_theta || !d_noise || !d_out || !d_info) return set_error(h, -2, "batched_lml_grad: null pointer");
    return launch_batched(h, d_X, d_Yc, d_theta, d_noise, nullptr, B, N, D, want_grad ? 1 : 0, d_out, d_info, nullptr, 0,
                          nullptr, nullptr);
}

int gpb_batched_lml_grad_ragged(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                                const double* d_noise, const int32_t* d_nrows, int64_t B, int64_t Nmax, int D, double* d_out,
                                int32_t* d_info, int want_grad) {
    GPB_ENTER(h);
    if (!d_X || !d_Yc || !d_theta || !d_noise || !d_nrows || !d_out || !d_info)
        return set_error(h, -2, "batched_lml_grad_ragged: null pointer");
    return launch_batched(h, d_X, d_Yc, d_theta, d_noise, d_nrows, B, Nmax, D, want_grad ? 1 : 0, d_out, d_info, nullptr, 0,
                          nullptr, nullptr);
}

int gpb_batched_predict_f_ragged(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                                 const double* d_noise, const int32_t* d_nrows, int64_t B, int64_t Nmax, int D,
                                 const double* d_Xs, int64_t Ns, double* d_mean, double* d_var, int32_t* d_info) {
    GPB_ENTER(h);
    if (!d_X || !d_Yc || !d_theta || !d_noise || !d_nrows || !d_Xs || !d_mean || !d_var || !d_info)
        return set_error(h, -2, "batched_predict_f_ragged: null pointer");
    if (Ns <= 0) return 0;
    return launch_batched(h, d_X, d_Yc, d_theta, d_noise, d_nrows, B, Nmax, D, 2, nullptr, d_info, d_Xs, Ns, d_mean, d_var);
}

int gpb_batched_predict_f(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                          const double* d_noise, int64_t B, int64_t N, int D, const double* d_Xs, int64_t Ns,
                          double* d_mean, double* d_var, int32_t* d_info) {
    GPB_ENTER(h);
    if (!d_X || !d_Yc || !d_theta || !d_noise || !d_Xs || !d_mean || !d_var || !d_info)
        return set_error(h, -2, "batched_predict_f: null pointer");
    if (Ns <= 0) return 0;
    return launch_batched(h, d_X, d_Yc, d_theta, d_noise, nullptr, B, N, D, 2, nullptr, d_info, d_Xs, Ns, d_mean, d_var);
}

}  // extern "C"

extern "C" {

int64_t gpb_svgp_flat_size(int64_t M, int D, int n_params) { return 2 + (int64_t)n_params + M * D + M + M * M; }

int gpb_svgp_data_term(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z, int64_t M, int D,
                       const double* d_qmu, const double* d_qsqrt, int64_t ldq, const double* d_Xb, const double* d_Yb,
                       int64_t B, double* d_flat, int want_grad) {
    GPB_ENTER(h);
    if (!h_theta || !d_Z || !d_qmu || !d_qsqrt || !d_Xb || !d_Yb || !d_flat) return set_error(h, -2, "svgp_data_term: null pointer");
    if (ldq < M) return set_error(h, -2, "svgp_data_term: ldq < M");
    return svgp_data_term(h, h_theta, noise_variance, d_Z, M, D, d_qmu, d_qsqrt, ldq, d_Xb, d_Yb, B, d_flat, want_grad);
}

int gpb_svgp_finish(gpb_handle* h, double* d_flat, double scale, const double* d_qmu, const double* d_qsqrt, int64_t ldq,
                    int64_t M, int D, int n_params, int apply_grad, double* h_elbo, double* h_kl) {
    GPB_ENTER(h);
    if (!d_flat || !d_qmu || !d_qsqrt || !h_elbo || !h_kl) return set_error(h, -2, "svgp_finish: null pointer");
    return svgp_finish(h, d_flat, scale, d_qmu, d_qsqrt, ldq, M, D, n_params, apply_grad, h_elbo, h_kl);
}

int gpb_svgp_predict_f(gpb_handle* h, const double* h_theta, const double* d_Z, int64_t M, int D, const double* d_qmu,
                       const double* d_qsqrt, int64_t ldq, const double* d_Xs, int64_t Ns, double* d_mean, double* d_var) {
    GPB_ENTER(h);
    if (!h_theta || !d_Z || !d_qmu || !d_qsqrt || !d_Xs || !d_mean || !d_var) return set_error(h, -2, "svgp_predict_f: null pointer");
    return svgp_predict_f(h, h_theta, d_Z, M, D, d_qmu, d_qsqrt, ldq, d_Xs, Ns, d_mean, d_var);
}

int gpb_sgpr_elbo(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z, int64_t M, int D,
                  const double* d_X, const double* d_err, int64_t N, int want_grad, double* h_out, double* d_errbar) {
    GPB_ENTER(h);
    if (!h_theta || !d_Z || !d_X || !d_err || !h_out) return set_error(h, -2, "sgpr_elbo: null pointer");
    if (M > 0x7fffffff || N > 0x7fffffff) return set_error(h, -2, "sgpr_elbo: size too large");
    return sgpr_elbo(h, h_theta, noise_variance, d_Z, M, D, d_X, d_err, N, want_grad, h_out, d_errbar);
}

int gpb_sgpr_predict_f(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z, int64_t M, int D,
                       const double* d_X, const double* d_err, int64_t N, const double* d_Xs, int64_t Ns, double* d_mean,
                       double* d_var) {
    GPB_ENTER(h);
    if (!h_theta || !d_Z || !d_X || !d_err) return set_error(h, -2, "sgpr_predict_f: null pointer");
    if (Ns > 0 && (!d_Xs || !d_mean || !d_var)) return set_error(h, -2, "sgpr_predict_f: null pointer");
    if (M > 0x7fffffff || N > 0x7fffffff || Ns > 0x7fffffff) return set_error(h, -2, "sgpr_predict_f: size too large");
    return sgpr_predict_f(h, h_theta, noise_variance, d_Z, M, D, d_X, d_err, N, d_Xs, Ns, d_mean, d_var);
}

}  // extern "C"
