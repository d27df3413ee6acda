// gpr.cu -- exact GP regression: log marginal likelihood, its hyper-parameter gradient, and
// predict_f, orchestrated over the fused assembly (assemble.cu), the DMMA GEMM (dgemm.cu) and the
// blocked factorisation (cholesky.cu).  North_star subsystems 2-3.
//
// Replaces gpflow/models/gpr.py GPR.log_marginal_likelihood + gpflow/logdensities.py
// multivariate_normal, the TF autodiff pass behind training_loss (K5) and
// gpflow/posteriors.py GPRPosterior / conditionals/util.py base_conditional_with_lm (G11).
// Reference call sites: GPR/model_trainer.py:18-20, GPR/predictor.py:6,
// Multi-Input_GPR/models/model_trainer.py:21,37, Multi-Input_GPR/main.py:434.
//
// Pipeline for one LML + gradient evaluation (all on the handle's stream, one sync at the end):
//   1. A  <- lower tiles of K(X,X) + noise I                      (assemble, HBM-bound)
//   2. A -> L blocks, W = L^-1                                   (factor_inv, 2N^3/3 flop on DMMA)
//   3. a = W y ; alpha = W^T a ; quad = a.a ; logdet = sum log L_ii
//   4. A  <- lower tiles of K^-1 = W^T W                          (lauum, N^3/3 flop on DMMA)
//   5. g_p = sum_{i>=j} c_ij (alpha_i alpha_j - K^-1_ij) dK_ij/dtheta_p   (fused, dK never stored)
//   6. LML = -quad/2 - N/2 log 2pi - logdet ; dLML/dtheta_p = g_p / 2 ; dLML/dnoise = tr(.)/2
#include <math.h>
#include <string.h>

#include "engine.cuh"

namespace gpb {

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct GprWork {
    double* A; double* W; int64_t ld;
    double* a; double* alpha; double* logdiag; double* res; int* info;
    double* tmp;            // N doubles: right-hand sides of the block substitutions (factor-only path)
    double* r; double* r2;  // 2 x N doubles: residual / correction of the refinement step (gpr_refine_alpha)
    double* Wd = nullptr;   // [N, NBD] strip of diagonal-block inverses (factor-only path, gpr_wd)
};

static int gpr_workspaces(gpb_handle* h, GprWork* w) {
    const int64_t N = h->N;
    w->ld = round_up(N, 16);
    const size_t mat = (size_t)round_up(N, 128) * w->ld * sizeof(double);
    w->A = workspace(h, BUF_K, mat);
    w->W = workspace(h, BUF_W, mat);
    const int64_t nblk = (N + 127) / 128;
    const size_t vec = (size_t)(6 * round_up(N, 16) + round_up(nblk, 16) + 64 + 16) * sizeof(double);
    double* v = workspace(h, BUF_VEC, vec);
    if (!w->A || !w->W || !v) return -1;
    w->a = v;
    w->alpha = v + round_up(N, 16);
    w->logdiag = w->alpha + round_up(N, 16);
    w->res = w->logdiag + round_up(nblk, 16);  // [0]=quad [1]=logdet [2..2+P]=grad sums, P+1 entries
    w->info = reinterpret_cast<int*>(w->res + 64);
    w->tmp = w->res + 64 + 16;
    w->r = w->tmp + round_up(N, 16);
    w->r2 = w->r + round_up(N, 16);
    return 0;
}

static int gpr_wd(gpb_handle* h, GprWork* w) {
    w->Wd = workspace(h, BUF_WD, (size_t)round_up(h->N, 128) * GPB_NBD * sizeof(double));
    return w->Wd ? 0 : -1;
}

// Is the factorisation in the workspaces the one for (current spec, theta, noise, bound X)?  The engine can
// only compare what it sees (pointer, sizes, parameters); that the CONTENT behind the X pointer is unchanged
// is the caller's guarantee, given by passing the serial it read after its own factorisation.
static bool factor_matches(const gpb_handle* h, bool was_valid, const double* theta, double noise, int64_t serial) {
    if (!was_valid || serial < 0 || serial != h->fact_serial) return false;
    if (h->fact_X != h->d_X || h->fact_N != h->N || h->fact_D != h->D || h->fact_noise != noise) return false;
    if (memcmp(&h->fact_spec, &h->spec, sizeof(gpb_kernel_spec)) != 0) return false;
    return memcmp(h->fact_theta, theta, sizeof(double) * (size_t)h->spec.n_params) == 0;
}

static void factor_remember(gpb_handle* h, const double* theta, double noise, int kind, bool has_alpha = true) {
    h->fact_valid = true;
    h->alpha_valid = has_alpha;
    h->fact_kind = kind;
    h->fact_serial += 1;
    h->fact_X = h->d_X; h->fact_N = h->N; h->fact_D = h->D; h->fact_noise = noise;
    h->fact_spec = h->spec;
    memset(h->fact_theta, 0, sizeof(h->fact_theta));
    memcpy(h->fact_theta, theta, sizeof(double) * (size_t)h->spec.n_params);
}

// steps 1-3; leaves W, a, alpha, res[0..1], info on the device.  reuse: W and the log-det partials of the
// previous call are still valid, only the vectors (which depend on the bound Y) are recomputed.
static int gpr_factor_matrix(gpb_handle* h, const DevKernel& kp, double noise, GprWork& w) {
    int rc;
    if ((rc = launch_assemble(h, kp, h->d_X, h->N, h->d_X, h->N, h->D, w.A, w.ld, 1, noise))) return rc;
    return factor_inv(h, w.A, w.ld, w.W, w.ld, h->N, w.logdiag, w.info, false);
}

// a = W y, alpha = W^T a, |a|^2 and the log-determinant: everything that depends on the bound targets
static int gpr_factor_vectors(gpb_handle* h, GprWork& w) {
    int rc;
    if ((rc = trmv_lower(h, w.W, w.ld, h->N, h->d_Yc, w.a))) return rc;
    if ((rc = trmv_lower_T(h, w.W, w.ld, h->N, w.a, w.alpha))) return rc;
    return quad_logdet(h, w.a, h->N, w.logdiag, w.res);
}

static int gpr_factor(gpb_handle* h, const DevKernel& kp, double noise, GprWork& w, bool reuse = false) {
    int rc;
    if (!reuse && (rc = gpr_factor_matrix(h, kp, noise, w))) return rc;
    return gpr_factor_vectors(h, w);
}

// Factor-only flow (no K^-1 wanted): L by factor_L (N^3/3 flop), a = L^-1 y by block forward substitution,
// |a|^2 and the log-determinant.  Leaves L (w.A diagonal blocks / w.W), w.Wd, w.a, res[0..1], info.
static int gpr_factor_only(gpb_handle* h, const DevKernel& kp, double noise, GprWork& w, bool reuse = false) {
    int rc;
    if (!reuse) {
        if ((rc = launch_assemble(h, kp, h->d_X, h->N, h->d_X, h->N, h->D, w.A, w.ld, 1, noise))) return rc;
        if ((rc = factor_L(h, w.A, w.ld, w.W, w.ld, w.Wd, h->N, w.logdiag, w.info))) return rc;
    }
    if ((rc = solve_L_vec(h, w.W, w.ld, w.Wd, h->N, h->d_Yc, w.a, w.tmp))) return rc;
    return quad_logdet(h, w.a, h->N, w.logdiag, w.res);
}

// One step of iterative refinement of alpha = (K + s2 I)^-1 y, for predict_f's mean = K(X*, X) alpha:
//   r = y - (K + s2 I) alpha   with K re-assembled from X row chunk by row chunk (never from the factor),
//   alpha += (L L^T)^-1 r      through whatever the handle holds (W = L^-1, or the factor + block inverses).
// Why: the engine solves by multiplying with explicit inverses; at the reference's own sigma^2 = 1e-5 on the
// C1 axis (cond 1.4e8) that left the predictive mean at 2.6e-7 of the extended-precision truth, LAPACK's
// triangular solves at 1.2e-8.  One step brings it to ~5e-9, the level of LAPACK's cho_solve (measured:
// tests/test_truth.py); cost one extra assembly pass and O(N^2) vector work.
// scratch: [rows, ldk] doubles for the row chunks of K.
static int gpr_refine_alpha(gpb_handle* h, const DevKernel& kp, double noise, GprWork& w, int kind, double* scratch,
                            int64_t rows, int64_t ldk) {
    const int64_t N = h->N;
    int rc;
    for (int64_t r0 = 0; r0 < N; r0 += rows) {
        const int64_t c = (N - r0 < rows) ? (N - r0) : rows;
        if ((rc = launch_assemble(h, kp, h->d_X + r0 * h->D, c, h->d_X, N, h->D, scratch, ldk, 0, 0.0))) return rc;
        if ((rc = gemv_sub(h, scratch, ldk, c, N, w.alpha, h->d_Yc + r0, noise, r0, w.r + r0))) return rc;
    }
    if (kind == 2) {
        if ((rc = solve_L_vec(h, w.W, w.ld, w.Wd, N, w.r, w.r2, w.tmp))) return rc;
        if ((rc = solve_LT_vec(h, w.W, w.ld, w.Wd, N, w.r2, w.r, w.tmp))) return rc;
    } else {
        if ((rc = trmv_lower(h, w.W, w.ld, N, w.r, w.tmp))) return rc;
        if ((rc = trmv_lower_T(h, w.W, w.ld, N, w.tmp, w.r))) return rc;
    }
    return vec_add(h, w.alpha, w.r, N);
}

// When is the objective itself refined?  cond(K + s2 I) <= (N max_x k(x, x) + s2) / s2 is known on the host:
// above ~2e7 the explicit-inverse solves leave |a|^2 = y^T K^-1 y at ~1e-9 relative (measured against the
// extended-precision truth on the C1 axis at the reference's sigma^2 = 1e-5: 6e-10 .. 1.1e-9; LAPACK 1e-11),
// so the quadratic form is then taken as y^T alpha with the refined alpha.  Below the threshold (every
// sigma^2 >= 1e-2 configuration) nothing is added to the evaluation.  gpb_set_option(h, 3, 0 / 1 / 2) =
// never / automatic / always.
static bool refine_objective(const gpb_handle* h, const DevKernel& kp, double noise) {
    if (h->refine_mode == 0) return false;
    if (h->refine_mode == 2) return true;
    double kb = 0.0;
    for (int t = 0; t < kp.n_terms; ++t) {
        double prod = 1.0;
        for (int f = 0; f < kp.terms[t].n_factors; ++f) {
            const DevLeaf& L = kp.leaves[kp.terms[t].leaf[f]];
            prod *= L.variance * (L.kind == GPB_LEAF_LINEAR ? 16.0 : 1.0);   // Linear: v |x|^2, z-scored inputs
        }
        kb += prod;
    }
    return (double)h->N * kb > 2e7 * noise;
}

// refined alpha (one step) and res[0] = y^T alpha; the caller has a, alpha (kind 1) or a (kind 2) in place
static int gpr_refine_objective(gpb_handle* h, const DevKernel& kp, double noise, GprWork& w, int kind) {
    const int64_t N = h->N, ldk = round_up(N, 16), rows = N < 1024 ? N : 1024;
    int rc;
    double* scratch = workspace(h, BUF_AUX2, (size_t)rows * ldk * sizeof(double));
    if (!scratch) return -1;
    if (kind == 2 && (rc = solve_LT_vec(h, w.W, w.ld, w.Wd, N, w.a, w.alpha, w.tmp))) return rc;
    if ((rc = gpr_refine_alpha(h, kp, noise, w, kind, scratch, rows, ldk))) return rc;
    return vec_dot(h, h->d_Yc, w.alpha, N, w.res);
}

int gpr_lml(gpb_handle* h, const double* theta, double noise, double* lml, double* grad_theta, double* grad_noise,
            int want_grad) {
    if (!h->has_spec) return set_error(h, -3, "gpr: no kernel set (gpb_set_kernel)");
    if (!h->d_X || h->N <= 0) return set_error(h, -3, "gpr: no data bound (gpb_gpr_set_data)");
    if (!(noise >= 0.0)) return set_error(h, -2, "gpr: noise variance must be >= 0");
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    if (kp.n_dims != h->D) return set_error(h, -2, "gpr: kernel expects D=%d, data has D=%d", kp.n_dims, h->D);
    GprWork w;
    if ((rc = gpr_workspaces(h, &w))) return rc;
    const int P = kp.n_params;
    const bool refine = refine_objective(h, kp, noise);
    if (want_grad && pipeline_applies(h, h->N) && h->side[0]) {
        // large N: right-looking pipeline over the two SM partitions; K^-1 accumulates row block by row block
        // inside it, the vector kernels start as soon as W is complete (beside the last K^-1 products)
        if ((rc = launch_assemble(h, kp, h->d_X, h->N, h->d_X, h->N, h->D, w.A, w.ld, 1, noise))) return rc;
        cudaEvent_t ev_W = nullptr;
        if ((rc = factor_inv_pipelined(h, w.A, w.ld, w.W, w.ld, h->N, w.logdiag, w.info, true, &ev_W))) return rc;
        cudaStream_t main_stream = h->stream, side = h->side[0];
        cudaError_t e = cudaStreamWaitEvent(side, ev_W, 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr pipeline fork");
        h->stream = side;
        rc = gpr_factor_vectors(h, w);
        if (!rc && refine) rc = gpr_refine_objective(h, kp, noise, w, 1);
        h->stream = main_stream;
        if (rc) return rc;
        e = cudaEventRecord(h->ev_join[0], side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, h->ev_join[0], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr pipeline join");
        factor_remember(h, theta, noise, 1);   // (a failed pivot invalidates it again below)
        if ((rc = launch_grad_reduce(h, kp, h->d_X, h->N, h->D, w.A, w.ld, w.alpha, w.res + 2))) return rc;
    } else if (want_grad && h->fork_streams && h->side[0]) {
        // the vector kernels (a, alpha, |a|^2, log-det) need W only, as does K^-1 = W^T W: they run on a side
        // stream beside the big product and meet again in front of the gradient reduction (which needs both)
        if ((rc = gpr_factor_matrix(h, kp, noise, w))) return rc;
        cudaStream_t main_stream = h->stream, side = h->side[0];
        cudaError_t e = cudaEventRecord(h->ev_fork[0], main_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, h->ev_fork[0], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr fork");
        h->stream = side;
        rc = gpr_factor_vectors(h, w);
        if (!rc && refine) rc = gpr_refine_objective(h, kp, noise, w, 1);
        h->stream = main_stream;
        if (rc) return rc;
        e = cudaEventRecord(h->ev_join[0], side);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr join record");
        if ((rc = lauum_lower(h, w.W, h->N, w.ld, w.A, w.ld))) return rc;
        e = cudaStreamWaitEvent(main_stream, h->ev_join[0], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr join");
        factor_remember(h, theta, noise, 1);   // (a failed pivot invalidates it again below)
        if ((rc = launch_grad_reduce(h, kp, h->d_X, h->N, h->D, w.A, w.ld, w.alpha, w.res + 2))) return rc;
    } else if (want_grad) {
        if ((rc = gpr_factor(h, kp, noise, w))) return rc;
        if (refine && (rc = gpr_refine_objective(h, kp, noise, w, 1))) return rc;
        factor_remember(h, theta, noise, 1);   // (a failed pivot invalidates it again below)
        if ((rc = lauum_lower(h, w.W, h->N, w.ld, w.A, w.ld))) return rc;
        if ((rc = launch_grad_reduce(h, kp, h->d_X, h->N, h->D, w.A, w.ld, w.alpha, w.res + 2))) return rc;
    } else {
        // value only: the factor alone (half the flop of factor + inverse)
        if ((rc = gpr_wd(h, &w))) return rc;
        if ((rc = gpr_factor_only(h, kp, noise, w))) return rc;
        if (refine && (rc = gpr_refine_objective(h, kp, noise, w, 2))) return rc;
        factor_remember(h, theta, noise, 2, refine);   // the value-only flow forms alpha only when it refines
    }
    // one D2H of [quad, logdet, g_0..g_P] + info, then the only sync of the evaluation
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    const size_t nres = (size_t)(2 + (want_grad ? P + 1 : 0));
    cudaError_t e = cudaMemcpyAsync(hp, w.res, nres * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hp + 64, w.info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "gpr result copy/sync");
    const int info = *reinterpret_cast<int*>(hp + 64);
    if (info > 0) {
        h->fact_valid = false;
        set_error(h, info, "Cholesky decomposition was not successful: non-positive pivot at row %d of %lld", info,
                  (long long)h->N);
        return info;
    }
    *lml = -0.5 * hp[0] - 0.5 * (double)h->N * log(2.0 * M_PI) - hp[1];
    if (want_grad) {
        for (int p = 0; p < P; ++p) grad_theta[p] = 0.5 * hp[2 + p];
        *grad_noise = 0.5 * hp[2 + P];
    }
    return 0;
}

int gpr_predict_f(gpb_handle* h, const double* theta, double noise, const double* d_Xs, int64_t Ns, double* d_mean,
                  double* d_var, int64_t reuse_serial) {
    if (!h->has_spec) return set_error(h, -3, "predict_f: no kernel set");
    if (!h->d_X || h->N <= 0) return set_error(h, -3, "predict_f: no data bound");
    if (Ns <= 0) return 0;
    DevKernel kp;
    int rc = build_dev_kernel(h, theta, &kp);
    if (rc) return rc;
    const bool was_valid = h->fact_valid;   // (the workspace requests below clear the flag)
    GprWork w;
    if ((rc = gpr_workspaces(h, &w))) return rc;
    const bool reuse = factor_matches(h, was_valid, theta, noise, reuse_serial);
    // a stored factorisation is used as it is (W = L^-1 after an objective + gradient evaluation, the factor
    // alone otherwise); a cold start takes the factor-only path: N^3/3 flop instead of 2N^3/3
    const int kind = reuse ? h->fact_kind : 2;
    if (kind == 2) {
        if ((rc = gpr_wd(h, &w))) return rc;
        if ((rc = gpr_factor_only(h, kp, noise, w, reuse))) return rc;
    } else {
        if ((rc = gpr_factor(h, kp, noise, w, true))) return rc;
    }
    if (reuse) h->fact_valid = true;        // same factorisation, same serial
    else factor_remember(h, theta, noise, 2);
    h->alpha_valid = true;                  // (both branches below leave alpha in place)
    const int64_t N = h->N;
    if (kind == 2 && (rc = solve_LT_vec(h, w.W, w.ld, w.Wd, N, w.a, w.alpha, w.tmp))) return rc;
    // chunk the test points so that the two [N, chunk] work matrices stay near 1 GiB each
    int64_t chunk = (int64_t)(1 << 27) / (N > 0 ? N : 1);
    chunk = chunk / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (chunk > Ns) chunk = round_up(Ns, 16);
    const int64_t ldc = round_up(chunk, 16);
    double* Kmn = workspace(h, BUF_AUX, (size_t)N * ldc * sizeof(double));
    // (Am doubles as the row-chunk scratch of the refinement step: at least 1024 rows of K(X, X))
    const int64_t ldk = round_up(N, 16);
    const int64_t am_doubles = (N * ldc > (N < 1024 ? N : 1024) * ldk) ? N * ldc : (N < 1024 ? N : 1024) * ldk;
    double* Am = workspace(h, BUF_AUX2, (size_t)am_doubles * sizeof(double));
    double* kd = workspace(h, BUF_PANEL, (size_t)ldc * sizeof(double));
    if (!Kmn || !Am || !kd) return -1;
    {
        // refinement scratch: the Am block seen as [rows, ldk] row chunks of K(X, X)
        int64_t rows = am_doubles / ldk;
        if (rows > N) rows = N;
        if (rows < 1) return set_error(h, -1, "predict_f: refinement scratch too small");
        if ((rc = gpr_refine_alpha(h, kp, noise, w, kind, Am, rows, ldk))) return rc;
    }
    for (int64_t s0 = 0; s0 < Ns; s0 += chunk) {
        const int64_t m = (Ns - s0 < chunk) ? (Ns - s0) : chunk;
        const double* Xs = d_Xs + s0 * h->D;
        if ((rc = launch_assemble(h, kp, h->d_X, N, Xs, m, h->D, Kmn, ldc, 0, 0.0))) return rc;
        if ((rc = launch_kdiag(h, kp, Xs, m, h->D, kd))) return rc;
        // mean = K(X*, X) alpha with the refined alpha (before the solve consumes Kmn)
        if ((rc = predict_colreduce(h, Kmn, ldc, N, m, w.alpha, nullptr, d_mean + s0, nullptr))) return rc;
        if (kind == 2) {
            // A = L^-1 Kmn by block forward substitution (Kmn is consumed)
            if ((rc = solve_L_mat(h, w.W, w.ld, w.Wd, N, Kmn, ldc, m, Am, ldc))) return rc;
        } else {
            GemmArgs g;  // A = W Kmn  (= L^-1 Kmn), W lower
            g.transa = 0; g.transb = 0; g.M = N; g.N = m; g.K = N;
            g.A = w.W; g.lda = w.ld; g.B = Kmn; g.ldb = ldc; g.C = Am; g.ldc = ldc; g.a_lower = 1;
            if ((rc = launch_gemm(h, g, h->stream))) return rc;
        }
        // var = kdiag - colsum(A^2)
        if ((rc = predict_colreduce(h, Am, ldc, N, m, w.a, kd, nullptr, d_var + s0))) return rc;
    }
    double* hp = pinned(h, (size_t)(64 + 2) * sizeof(double));
    if (!hp) return -1;
    cudaError_t e = cudaMemcpyAsync(hp + 64, w.info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "predict_f sync");
    const int info = *reinterpret_cast<int*>(hp + 64);
    if (info > 0) {
        h->fact_valid = false;
        set_error(h, info, "Cholesky decomposition was not successful: non-positive pivot at row %d of %lld", info,
                  (long long)N);
        return info;
    }
    return 0;
}

}  // namespace gpb

namespace gpb {

// alpha = (K + noise I)^-1 (y - m(X)) of the LAST gpr_lml / gpr_predict_f evaluation on this handle
// (= d LML / d m(X): what the host layer needs to train mean-function parameters).
int gpr_get_alpha(gpb_handle* h, double* d_alpha) {
    if (!h->d_X || h->N <= 0) return set_error(h, -3, "get_alpha: no data bound");
    if (!h->fact_valid || !h->alpha_valid)
        return set_error(h, -3, "get_alpha: no alpha held (call gpb_gpr_lml_grad or gpb_gpr_predict_f on this handle first)");
    GprWork w;
    const bool keep = h->fact_valid;   // read-only use of the workspaces
    int rc = gpr_workspaces(h, &w);
    h->fact_valid = keep;
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(d_alpha, w.alpha, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
    return check_cuda(h, e, "get_alpha copy");
}

}  // namespace gpb

extern "C" int gpb_gpr_get_alpha(gpb_handle* h, double* d_alpha) {
    GPB_ENTER(h);
    if (!d_alpha) return gpb::set_error(h, -2, "get_alpha: null pointer");
    return gpb::gpr_get_alpha(h, d_alpha);
}
