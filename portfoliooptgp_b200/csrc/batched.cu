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
#include "batched_kernel.cuh"

namespace gpb {

int launch_batched(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta, const double* d_noise,
                   const int* d_nrows, int64_t B, int64_t N, int D, int mode, double* d_out, int* d_info, const double* d_Xs,
                   int64_t Ns, double* d_mean, double* d_var) {
    if (!h->has_spec) return set_error(h, -3, "batched: no kernel set (gpb_set_kernel)");
    if (B <= 0) return 0;
    if (N < 1 || N > 128) return set_error(h, -2, "batched: N=%lld outside [1,128] (one GP per CTA in shared memory)", (long long)N);
    if (D != h->spec.n_dims) return set_error(h, -2, "batched: kernel expects D=%d, got %d", h->spec.n_dims, D);
    if (B > 0x7fffffffLL) return set_error(h, -2, "batched: B too large");
    int dp = 1;
    while (dp < D) dp <<= 1;
    bool fast = (h->spec.n_leaves <= GRAD_FAST_LEAVES);
    for (int g = 0; g < h->spec.n_groups; ++g)
        if (h->spec.groups[g].ard_index >= 0) fast = false;
    const int n = (int)N, ns = (int)Ns;
    if (!fast) return launch_batched_generic(dp, h, d_X, d_Yc, d_theta, d_noise, d_nrows, B, n, D, mode, d_out, d_info, d_Xs, ns, d_mean, d_var);
    // expression shape (structure only: a descriptor built with a dummy theta is enough to match it)
    if (h->use_shapes) {
        double ones[GPB_MAX_PARAMS];
        for (int p = 0; p < GPB_MAX_PARAMS; ++p) ones[p] = 1.0;
        DevKernel probe;
        build_dev_kernel_core(h->spec, ones, &probe);
        const int shape = match_shape(probe);
        if (shape != SHAPE_NONE) {
            const int rc = launch_batched_static(shape, dp, h, d_X, d_Yc, d_theta, d_noise, d_nrows, B, n, D, mode, d_out, d_info,
                                                 d_Xs, ns, d_mean, d_var);
            if (rc != -100) return rc;   // -100: no straight-line instantiation for this (shape, DP)
        }
    }
    {
        const int N = n, Ns = ns;   // (shadow the 64-bit arguments for the launcher macro)
        switch (dp) {
            case 1: return launch_batched_dp<1, true>(GPB_BATCHED_ARGS);
            case 2: return launch_batched_dp<2, true>(GPB_BATCHED_ARGS);
            case 4: return launch_batched_dp<4, true>(GPB_BATCHED_ARGS);
            case 8: return launch_batched_dp<8, true>(GPB_BATCHED_ARGS);
            default: return launch_batched_dp<16, true>(GPB_BATCHED_ARGS);
        }
    }
}

}  // namespace gpb
