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
};

static int gpr_workspaces(gpb_handle* h, GprWork* w) {
    const int64_t N = h->N;
    w->ld = round_up(N, 16);
    const size_t mat = (size_t)round_up(N, 128) * w->ld * sizeof(double);
    w->A = workspace(h, BUF_K, mat);
    w->W = workspace(h, BUF_W, mat);
    const int64_t nblk = (N + 127) / 128;
    const size_t vec = (size_t)(2 * round_up(N, 16) + round_up(nblk, 16) + 64 + 16) * sizeof(double);
    double* v = workspace(h, BUF_VEC, vec);
    if (!w->A || !w->W || !v) return -1;
    w->a = v;
    w->alpha = v + round_up(N, 16);
    w->logdiag = w->alpha + round_up(N, 16);
    w->res = w->logdiag + round_up(nblk, 16);  // [0]=quad [1]=logdet [2..2+P]=grad sums, P+1 entries
    w->info = reinterpret_cast<int*>(w->res + 64);
    return 0;
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

static void factor_remember(gpb_handle* h, const double* theta, double noise) {
    h->fact_valid = true;
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
    if (want_grad && h->fork_streams && h->side[0]) {
        // the vector kernels (a, alpha, |a|^2, log-det) need W only, as does K^-1 = W^T W: they run on a side
        // stream beside the big product and meet again in front of the gradient reduction (which needs both)
        if ((rc = gpr_factor_matrix(h, kp, noise, w))) return rc;
        cudaStream_t main_stream = h->stream, side = h->side[0];
        cudaError_t e = cudaEventRecord(h->ev_fork[0], main_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, h->ev_fork[0], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr fork");
        h->stream = side;
        rc = gpr_factor_vectors(h, w);
        h->stream = main_stream;
        if (rc) return rc;
        e = cudaEventRecord(h->ev_join[0], side);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr join record");
        if ((rc = lauum_lower(h, w.W, h->N, w.ld, w.A, w.ld))) return rc;
        e = cudaStreamWaitEvent(main_stream, h->ev_join[0], 0);
        if (e != cudaSuccess) return check_cuda(h, e, "gpr join");
        factor_remember(h, theta, noise);   // (a failed pivot invalidates it again below)
        if ((rc = launch_grad_reduce(h, kp, h->d_X, h->N, h->D, w.A, w.ld, w.alpha, w.res + 2))) return rc;
    } else {
        if ((rc = gpr_factor(h, kp, noise, w))) return rc;
        factor_remember(h, theta, noise);   // (a failed pivot invalidates it again below)
        if (want_grad) {
            if ((rc = lauum_lower(h, w.W, h->N, w.ld, w.A, w.ld))) return rc;
            if ((rc = launch_grad_reduce(h, kp, h->d_X, h->N, h->D, w.A, w.ld, w.alpha, w.res + 2))) return rc;
        }
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
    if ((rc = gpr_factor(h, kp, noise, w, reuse))) return rc;
    if (reuse) h->fact_valid = true;        // same factorisation, same serial
    else factor_remember(h, theta, noise);
    const int64_t N = h->N;
    // chunk the test points so that the two [N, chunk] work matrices stay near 1 GiB each
    int64_t chunk = (int64_t)(1 << 27) / (N > 0 ? N : 1);
    chunk = chunk / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (chunk > Ns) chunk = round_up(Ns, 16);
    const int64_t ldc = round_up(chunk, 16);
    double* Kmn = workspace(h, BUF_AUX, (size_t)N * ldc * sizeof(double));
    double* Am = workspace(h, BUF_AUX2, (size_t)N * ldc * sizeof(double));
    double* kd = workspace(h, BUF_PANEL, (size_t)ldc * sizeof(double));
    if (!Kmn || !Am || !kd) return -1;
    for (int64_t s0 = 0; s0 < Ns; s0 += chunk) {
        const int64_t m = (Ns - s0 < chunk) ? (Ns - s0) : chunk;
        const double* Xs = d_Xs + s0 * h->D;
        if ((rc = launch_assemble(h, kp, h->d_X, N, Xs, m, h->D, Kmn, ldc, 0, 0.0))) return rc;
        if ((rc = launch_kdiag(h, kp, Xs, m, h->D, kd))) return rc;
        GemmArgs g;  // A = W Kmn  (= L^-1 Kmn), W lower
        g.transa = 0; g.transb = 0; g.M = N; g.N = m; g.K = N;
        g.A = w.W; g.lda = w.ld; g.B = Kmn; g.ldb = ldc; g.C = Am; g.ldc = ldc; g.a_lower = 1;
        if ((rc = launch_gemm(h, g, h->stream))) return rc;
        // var = kdiag - colsum(A^2) ; mean = A^T (L^-1 y)
        if ((rc = predict_colreduce(h, Am, ldc, N, m, w.a, kd, d_mean + s0, d_var + s0))) return rc;
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
