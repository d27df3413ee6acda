/* gpb200.h -- C-ABI of the B200-native exact-GP / SVGP engine (libgpb200.so).
 *
 * Drop-in boundary for the GPflow model path that PortfolioOptGP drives.
 * The reference has no FFI of its own (it is pure Python on GPflow 2.9.1 / TensorFlow-CPU); every
 * entry point below therefore cites the reference call site and the GPflow routine whose arithmetic
 * it replaces.  Paths are relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices fp64, row-major, leading dimension in elements;
 *   - pointers named d_* are DEVICE pointers on the handle's device (borrowed for the call, or for
 *     the lifetime of the binding for gpb_gpr_set_data); h_* are HOST pointers;
 *   - every call returns 0 on success, >0 = LAPACK-style index (1-based) of the first non-positive
 *     Cholesky pivot (GPflow/TF: InvalidArgumentError "Cholesky decomposition was not
 *     successful"), <0 = bad argument / CUDA error, text via gpb_last_error();
 *   - work is enqueued on the handle's stream; calls that return host scalars synchronise that
 *     stream before returning, the others are asynchronous;
 *   - a handle is bound to one device and is not thread-safe; distinct handles are independent.
 *   - hyper-parameters cross the boundary in CONSTRAINED space (variance, lengthscale, ...); the
 *     softplus chain rule of gpflow.Parameter stays in the host layer.
 */
#ifndef GPB200_H
#define GPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_MAX_DIMS 16    /* input columns D the fused assembly supports */
#define GPB_MAX_GROUPS 8   /* distinct distance computations per kernel expression */
#define GPB_MAX_LEAVES 8   /* leaf kernels per expression */
#define GPB_MAX_TERMS 8    /* products in the sum-of-products normal form */
#define GPB_MAX_FACTORS 4  /* leaves per product */
#define GPB_MAX_PARAMS 48  /* flat constrained hyper-parameter vector length */

/* How a group reduces a pair (x, x') over its active dims to one scalar s. */
enum gpb_group_kind {
    GPB_GROUP_EUCLID = 0,       /* s = sum_d w_d (x_d - x'_d)^2      gpflow/utilities/ops.py square_distance */
    GPB_GROUP_PERIODIC_SQ = 1,  /* s = sum_d w_d sin^2(pi (x_d - x'_d)/p)   gpflow/kernels/periodic.py, K_r2 bases */
    GPB_GROUP_PERIODIC_ABS = 2, /* s = sum_d w_d |sin(pi (x_d - x'_d)/p)|   gpflow/kernels/periodic.py, K_r bases */
    GPB_GROUP_DOT = 3           /* s = sum_d w_d x_d x'_d            gpflow/kernels/linears.py Linear */
};

/* gpflow/kernels/stationaries.py + linears.py leaf formulas. */
enum gpb_leaf_kind {
    GPB_LEAF_SE = 0,          /* variance * exp(-r2/2) */
    GPB_LEAF_RQ = 1,          /* variance * (1 + r2/(2 alpha))^-alpha */
    GPB_LEAF_MATERN12 = 2,    /* variance * exp(-r) */
    GPB_LEAF_EXPONENTIAL = 3, /* variance * exp(-r/2) */
    GPB_LEAF_MATERN32 = 4,
    GPB_LEAF_MATERN52 = 5,
    GPB_LEAF_LINEAR = 6       /* variance * s */
};

typedef struct {
    int32_t kind;         /* gpb_group_kind */
    uint32_t dim_mask;    /* bit d set <=> column d is active (GPflow active_dims) */
    int32_t ard_index;    /* >=0: theta[ard_index + k] is the lengthscale of the k-th active dim (ARD); -1: isotropic */
    int32_t period_index; /* periodic groups: theta index of the (scalar) period; else -1 */
} gpb_group;

typedef struct {
    int32_t kind;        /* gpb_leaf_kind */
    int32_t group;       /* index into groups[] */
    int32_t var_index;   /* theta index of variance */
    int32_t ls_index;    /* theta index of the scalar lengthscale; -1 for ARD (see group) and Linear */
    int32_t alpha_index; /* RationalQuadratic alpha; else -1 */
} gpb_leaf;

typedef struct {
    int32_t n_factors;
    int32_t leaf[GPB_MAX_FACTORS];
} gpb_term;

/* Kernel expression in sum-of-products normal form: K = sum_t prod_f leaf[t][f].
 * Replaces gpflow.kernels.{Sum,Product} trees built at GPR/main.py:105-114 and
 * Multi-Input_GPR/main.py:126-135 (k1(active_dims) * k2(active_dims)). */
typedef struct {
    int32_t n_dims;   /* D of the X this expression is applied to */
    int32_t n_params; /* length of the flat constrained theta vector */
    int32_t n_groups, n_leaves, n_terms;
    gpb_group groups[GPB_MAX_GROUPS];
    gpb_leaf leaves[GPB_MAX_LEAVES];
    gpb_term terms[GPB_MAX_TERMS];
} gpb_kernel_spec;

typedef struct gpb_handle gpb_handle;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int gpb_create(gpb_handle** out, int device);
int gpb_destroy(gpb_handle* h);
const char* gpb_last_error(gpb_handle* h);
/* cuda_stream is a cudaStream_t (NULL = legacy default stream).  Lets the torch host layer order
 * engine work against its own allocations. */
int gpb_set_stream(gpb_handle* h, void* cuda_stream);
/* ABI version and build info ("sm_100a ..."); */
int gpb_version(void);
/* Number of engine kernels launched through this handle since creation (bench.py gpu_launches). */
int64_t gpb_launch_count(gpb_handle* h);

/* Engine options.  option 0: fork off-critical-path products of the blocked factorisation onto side
 * streams (default 1; bench.py switches it off while it times individual kernels).  option 1: launch
 * the GEMM / leaf kernels with programmatic dependent launch (default 1).  option 2: use the
 * straight-line instantiations of the element kernels for the known expression shapes
 * (csrc/shapes.cuh; default 1, 0 forces the run-time interpreter -- same results).  option 3: one
 * step of iterative refinement of alpha = (K + s2 I)^-1 y for the quadratic form of the objective
 * (0 never, 1 automatic: only when the host-side bound N k(x,x) / s2 on cond(K + s2 I) exceeds 2e7, 2
 * always; default 1).  predict_f always refines alpha for the mean (csrc/gpr.cu).  option 4: the
 * pipelined right-looking factorisation over two SM partitions (green contexts; csrc/partition.cu,
 * csrc/cholesky.cu) for N >= 3072 (default 0 = the single-partition recursion: on B200 the pipeline's
 * rank-1024 products run at 0.90 of the recursion's large-K products and the 8 reserved SMs cost 5 %,
 * which outweighs the hidden latency -- 22.8 vs 21.3 ms at N = 8192, profiles/r02_pipeline_ab.txt;
 * silently off when the driver cannot create the partitions).  option 5: the look-ahead chain over the
 * 128-row leaves of every <= 1024-row (2048 inside matrices of >= 4096 rows) diagonal block at the bottom of the blocked factorisation
 * (default 1; 0 = the plain 2 x 2 recursion down to the leaves). */
int gpb_set_option(gpb_handle* h, int option, int value);
/* Diagnostics: which straight-line shape (csrc/shapes.cuh, 1-based id) the current expression matches;
 * 0 = none, the run-time interpreter evaluates it (also when option 2 is off).  < 0: error. */
int gpb_kernel_shape(gpb_handle* h);

/* Optional kernel timing with CUDA events recorded on the handle's stream around each engine
 * launch, by category (0 DMMA GEMM, 1 assembly, 2 Cholesky leaf, 3 fused gradient reduction,
 * 4 vector kernels, 5 batched, 6 SVGP, 7 DMMA GEMM in its small-tile latency-bound configurations;
 * category 0 holds the large-tile launches only).  gpb_profile_read synchronises, returns the summed
 * milliseconds and launch counts per category (arrays of 8) and resets the log;
 * gpb_profile_read_flops returns (and resets) the flop those GEMM launches executed, by category.
 * bench.py's roofline numbers come from here; leave it off when timing end to end. */
int gpb_profile_enable(gpb_handle* h, int on);
int gpb_profile_read(gpb_handle* h, double* h_ms, int64_t* h_counts);
int gpb_profile_read_flops(gpb_handle* h, double* h_flops);

/* ---- kernel expression ---------------------------------------------------------------------- */
/* Replaces the kernel object handed to gpflow.models.GPR(kernel=...) (GPR/model_trainer.py:15). */
int gpb_set_kernel(gpb_handle* h, const gpb_kernel_spec* spec);

/* ---- fused kernel-matrix assembly (north_star subsystem 1) --------------------------------------
 * Replaces gpflow Kernel.__call__ -> K / K_diag (+ GPR._add_noise_cov, Kuu jitter) i.e. TF MatMul +
 * Exp/Sqrt/Sin + AddN/Mul + set_diag, for GPR/model_trainer.py:15-20 and GPR/predictor.py:6.
 * mode: 0 = full [N,N2] (d_X2 may differ from d_X); 1 = lower triangle of K(X,X) only (tiles on or
 * below the diagonal; what the Cholesky consumes); 2 = K(X,X) computing lower tiles once and
 * mirroring them (symmetric full).  diag_add is added to K[i][i] in modes 1,2 (noise / jitter). */
int gpb_assemble(gpb_handle* h, const double* h_theta, const double* d_X, int64_t N, const double* d_X2,
                 int64_t N2, int D, double* d_K, int64_t ldk, int mode, double diag_add);
int gpb_kdiag(gpb_handle* h, const double* h_theta, const double* d_X, int64_t N, int D, double* d_out);

/* ---- dense fp64 building blocks (north_star subsystem 2), exposed for tests / reuse ------------
 * Replace tf.linalg.cholesky / tf.linalg.triangular_solve / tf.matmul. */
/* In-place lower Cholesky of the row-major lower triangle of d_A [N,N] (tf.linalg.cholesky).  The
 * strict upper triangle is workspace: zero inside 128-aligned diagonal blocks, unspecified elsewhere. */
int gpb_potrf(gpb_handle* h, double* d_A, int64_t N, int64_t lda);
/* As gpb_potrf, and additionally d_W (lower) = L^-1 (what tf.linalg.triangular_solve(L, .) applies). */
int gpb_potrf_inv(gpb_handle* h, double* d_A, int64_t N, int64_t lda, double* d_W, int64_t ldw);
/* d_Out (lower tiles) = d_W^T d_W, i.e. (L L^T)^-1 when d_W = L^-1 from gpb_potrf_inv. */
int gpb_lauum(gpb_handle* h, const double* d_W, int64_t N, int64_t ldw, double* d_Out, int64_t ldo);
/* C = alpha * op(A) op(B) + beta * C (tf.matmul); transa/transb: 0 = as stored, 1 = transposed.
 * Row-major.  tri: 0 = full C, 1 = only tiles of C on/below the diagonal are computed (SYRK-style). */
int gpb_gemm(gpb_handle* h, int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha,
             const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double beta, double* d_C,
             int64_t ldc, int tri);

/* ---- exact GP regression (north_star subsystems 2-3) --------------------------------------------
 * Replaces gpflow.models.GPR((X, Y), kernel, noise_variance) at GPR/model_trainer.py:15,
 * Multi-Input_GPR/main.py:421-423, models/model_trainer.py:31.  d_X [N,D] row-major (ld = D),
 * d_Yc [N] = Y - mean_function(X) (R = 1 on every reference call site).  Buffers are borrowed until
 * the next set_data / destroy. */
int gpb_gpr_set_data(gpb_handle* h, const double* d_X, int64_t N, int D, const double* d_Yc);
/* log marginal likelihood: GPR.log_marginal_likelihood (gpflow/models/gpr.py) =
 * -1/2 |L^-1 y|^2 - N/2 log 2pi - sum log L_ii, with L = chol(K + noise I). */
int gpb_gpr_lml(gpb_handle* h, const double* h_theta, double noise_variance, double* h_lml);
/* LML and d LML / d theta (constrained, n_params entries) and d LML / d noise_variance, via
 * 1/2 tr((a a^T - K^-1) dK/dtheta) with dK/dtheta recomputed on the fly -- replaces the
 * tf.GradientTape pass inside gpflow.optimizers.Scipy.minimize (GPR/model_trainer.py:18-19). */
int gpb_gpr_lml_grad(gpb_handle* h, const double* h_theta, double noise_variance, double* h_lml,
                     double* h_grad_theta, double* h_grad_noise);
/* GPR.predict_f(Xnew, full_cov=False) (gpflow/posteriors.py GPRPosterior, conditionals/util.py
 * base_conditional_with_lm), GPR/predictor.py:6, Multi-Input_GPR/main.py:434.
 * d_mean, d_var: [Ns] device outputs (mean excludes mean_function(Xnew)). */
int gpb_gpr_predict_f(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Xs,
                      int64_t Ns, double* d_mean, double* d_var);

/* predict_f that may skip the factorisation.  gpb_gpr_factor_serial(h) identifies the factorisation the
 * last gpb_gpr_lml / gpb_gpr_lml_grad / gpb_gpr_predict_f call left in the handle's workspaces.  A caller
 * that reads it right after its own call and passes it back here gets W = L^-1 reused when the kernel
 * expression, theta, noise and the bound X (pointer, N, D) are bit-identical and nothing else has used the
 * workspaces in between; otherwise the call silently does the full work (same results either way).  The
 * engine cannot see the memory behind the X pointer: that its CONTENT is unchanged is the caller's
 * guarantee (the host layer keeps the tensor alive in the model object).  Use: GPflow's predict_y
 * recomputes predict_f, and the reference calls both back to back (GPR/predictor.py:6-7). */
int64_t gpb_gpr_factor_serial(gpb_handle* h);
int gpb_gpr_predict_f_reuse(gpb_handle* h, const double* h_theta, double noise_variance,
                            int64_t factor_serial, const double* d_Xs, int64_t Ns, double* d_mean,
                            double* d_var);

/* MANY independent exact GPs of any size in flight on one GPU: njobs evaluations (gpb_gpr_set_data +
 * gpb_gpr_lml_grad, or gpb_gpr_lml when want_grad == 0) distributed over nh handles of the same device, one
 * host thread per handle (job j runs on handle j % nh; the handles need their own streams and the same kernel
 * expression set).  A single evaluation at N = 129 .. ~2000 is bound by the one-CTA chain of its blocked
 * factorisation and leaves the machine idle; nh of them side by side fill it.  This is the rolling re-fit of
 * Multi-Input_GPR/main.py:414-456 for windows longer than the 128 rows of the one-GP-per-CTA path (the loop's
 * windows grow by one row per test day), and the restart / candidate loops of models/model_trainer.py:26-48,
 * GPR/model_trainer.py:10-26 without a host thread of the caller per fit.
 * d_X[j] [N[j], D], d_Yc[j] [N[j]] device pointers; h_theta [njobs, P] row-major, h_noise [njobs];
 * outputs h_lml [njobs], h_grad_theta [njobs, P], h_grad_noise [njobs] (ignored when want_grad == 0),
 * h_rc [njobs] = the return code of job j (0 ok, > 0 first non-positive pivot, < 0 error: gpb_last_error of
 * handle j % nh).  Returns 0 when every job ran (whatever its own code), < 0 on bad arguments. */
int gpb_gpr_lml_grad_many(int nh, gpb_handle* const* handles, int64_t njobs, const double* const* d_X,
                          const int64_t* N, int D, const double* const* d_Yc, const double* h_theta, int P,
                          const double* h_noise, int want_grad, double* h_lml, double* h_grad_theta,
                          double* h_grad_noise, int* h_rc);

/* The same for predict_f: job j = gpb_gpr_set_data(d_X[j], N[j], D, d_Yc[j]) + gpb_gpr_predict_f(theta[j], noise[j],
 * d_Xs[j] [Ns, D]) -> d_mean[j], d_var[j] [Ns] on handle j % nh (Multi-Input_GPR/main.py:434,454: the one-step-ahead
 * prediction of every re-fitted window).  h_rc [njobs] as above; outputs of a job with h_rc != 0 are undefined.
 * Every handle's stream is synchronised before the call returns. */
int gpb_gpr_predict_f_many(int nh, gpb_handle* const* handles, int64_t njobs, const double* const* d_X,
                           const int64_t* N, int D, const double* const* d_Yc, const double* h_theta, int P,
                           const double* h_noise, const double* const* d_Xs, int64_t Ns, double* const* d_mean,
                           double* const* d_var, int* h_rc);

/* d_alpha [N] <- (K + noise I)^-1 (Y - m(X)) of the last gpb_gpr_lml / gpb_gpr_lml_grad / gpb_gpr_predict_f
 * evaluation on this handle (= dLML/dm(X); lets the host layer train mean-function parameters,
 * test_scripts/GPFlow.py:186-190 uses Constant / Linear mean functions).  Asynchronous. */
int gpb_gpr_get_alpha(gpb_handle* h, double* d_alpha);

/* ---- batched independent small GPs (north_star subsystem 4) -------------------------------------
 * One GP per CTA, K resident in shared memory (N <= 128).  Replaces the sequential rolling re-fit
 * loop Multi-Input_GPR/main.py:414-456 x restarts models/model_trainer.py:26-48: B independent
 * GPR objective(+gradient) evaluations sharing one kernel expression (gpb_set_kernel).
 * All pointers are DEVICE pointers: d_X [B,N,D], d_Yc [B,N], d_theta [B,n_params] (constrained),
 * d_noise [B]; d_out [B, 2 + n_params] = {lml, dlml/dnoise, dlml/dtheta...} (gradient entries are
 * written only when want_grad); d_info [B]: 0 ok, >0 first non-positive pivot, <0 bad lengthscale.
 * Asynchronous on the handle's stream. */
int gpb_batched_lml_grad(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                         const double* d_noise, int64_t B, int64_t N, int D, double* d_out,
                         int32_t* d_info, int want_grad);
/* Batched GPR.predict_f(full_cov=False) at Ns new points per GP (Multi-Input_GPR/main.py:434 takes
 * the last row): d_Xs [B,Ns,D] -> d_mean, d_var [B,Ns]. */
int gpb_batched_predict_f(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                          const double* d_noise, int64_t B, int64_t N, int D, const double* d_Xs,
                          int64_t Ns, double* d_mean, double* d_var, int32_t* d_info);
/* Ragged variants: GP b uses only the first d_nrows[b] (1 <= d_nrows[b] <= Nmax) of its Nmax rows of
 * d_X [B,Nmax,D] / d_Yc [B,Nmax] -- the reference's EXPANDING windows X_full[:i], i = i0, i0+1, ...
 * (Multi-Input_GPR/main.py:414-423) as one batch, each GP with its own N.  Everything else as above. */
int gpb_batched_lml_grad_ragged(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                                const double* d_noise, const int32_t* d_nrows, int64_t B, int64_t Nmax, int D,
                                double* d_out, int32_t* d_info, int want_grad);
int gpb_batched_predict_f_ragged(gpb_handle* h, const double* d_X, const double* d_Yc, const double* d_theta,
                                 const double* d_noise, const int32_t* d_nrows, int64_t B, int64_t Nmax, int D,
                                 const double* d_Xs, int64_t Ns, double* d_mean, double* d_var, int32_t* d_info);

/* ---- SVGP (north_star subsystem 5) ---------------------------------------------------------------
 * Replaces gpflow.models.SVGP(kernel, Gaussian, Z, num_data).elbo / training_loss_closure /
 * predict_f (test_scripts/SVGP.py:515-540): whiten=True, q_diag=False, one latent GP.
 * All d_* are device pointers: d_Z [M,D], d_qmu [M], d_qsqrt [M, ldq] lower triangular with zeros
 * above the diagonal, d_Xb [B,D], d_Yb [B].
 *
 * gpb_svgp_data_term: minibatch pass.  Writes the flat record d_flat (gpb_svgp_flat_size doubles):
 *   [0] S = sum_b E_q[log p(y_b|f_b)]  (UNSCALED)   [1] dS/dnoise   [2,2+P) dS/dtheta
 *   then dS/dZ [M,D], dS/dq_mu [M], dS/dq_sqrt [M,M] (row-major).  Asynchronous.  Data-parallel
 *   ranks sum their records with ONE all-reduce (NCCL) before gpb_svgp_finish.
 * gpb_svgp_finish: ELBO = scale * S - KL[q(u)||p(u)]; with apply_grad the record is turned in place
 *   into d ELBO / d(noise, theta, Z, q_mu, q_sqrt) (KL added once -- SURVEY.md H8).  Synchronises. */
int64_t gpb_svgp_flat_size(int64_t M, int D, int n_params);
int gpb_svgp_data_term(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z,
                       int64_t M, int D, const double* d_qmu, const double* d_qsqrt, int64_t ldq,
                       const double* d_Xb, const double* d_Yb, int64_t B, double* d_flat, int want_grad);
int gpb_svgp_finish(gpb_handle* h, double* d_flat, double scale, const double* d_qmu,
                    const double* d_qsqrt, int64_t ldq, int64_t M, int D, int n_params, int apply_grad,
                    double* h_elbo, double* h_kl);
/* SVGP.predict_f(Xnew, full_cov=False): d_mean, d_var [Ns] (mean excludes mean_function). */
int gpb_svgp_predict_f(gpb_handle* h, const double* h_theta, const double* d_Z, int64_t M, int D,
                       const double* d_qmu, const double* d_qsqrt, int64_t ldq, const double* d_Xs,
                       int64_t Ns, double* d_mean, double* d_var);

/* ---- SGPR: Titsias' collapsed sparse bound (gpflow/models/sgpr.py SGPR.elbo / predict_f; the model
 * built at test_scripts/SVGP.py:393-399 and trained with Scipy).  d_Z [M, D] inducing points, d_X [N, D],
 * d_err [N] = Y - mean_function(X); jitter 1e-6 on Kuu as gpflow.config.default_jitter().
 * h_out (host) [2 + n_params + M*D] = elbo, d elbo/d noise_variance, d elbo/d theta (constrained, engine
 * order), d elbo/d Z (row-major); only h_out[0] is written when want_grad == 0.  d_errbar [N] (device,
 * may be NULL) = d elbo/d err for mean-function parameters.  Returns > 0 = 1-based failing pivot of
 * Kuu or of I + A A^T ("Cholesky decomposition was not successful"). */
int gpb_sgpr_elbo(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z, int64_t M,
                  int D, const double* d_X, const double* d_err, int64_t N, int want_grad, double* h_out,
                  double* d_errbar);
/* SGPR.predict_f(Xnew, full_cov=False): d_mean, d_var [Ns] (mean excludes mean_function(Xnew)). */
int gpb_sgpr_predict_f(gpb_handle* h, const double* h_theta, double noise_variance, const double* d_Z,
                       int64_t M, int D, const double* d_X, const double* d_err, int64_t N,
                       const double* d_Xs, int64_t Ns, double* d_mean, double* d_var);

/* Adam update of a device-resident parameter block in place (x += / -= lr * mhat / (sqrt(vhat) + eps));
 * used by the data-parallel minibatch SVGP loop for Z, q_mu, q_sqrt (identity transforms), where
 * GPflow users run tf.optimizers.Adam on minibatches.  step >= 1 is the 1-based iteration. */
int gpb_adam_step(gpb_handle* h, double* d_x, const double* d_g, double* d_m, double* d_v, int64_t n,
                  double lr, double beta1, double beta2, double eps, int64_t step, int maximize);

/* ---- device-side data preparation (the step before the path) ---------------
 * Series are [T, A] row-major in device memory (T time steps, one column per asset / field).
 * Replaces the pandas arithmetic of Multi-Input_GPR/utils/data_handler.py:86-91 (returns),
 * :160-169 (normalize_and_reshape), :129-154 (concatenate_X) and the window slicing of
 * Multi-Input_GPR/main.py:414-423 for data that is already resident in HBM.
 *
 * gpb_prep_returns: kind 0 = close.pct_change() with row 0 filled from row 1 (data_handler.py:86-88);
 *   kind 1 = (close - open) / open (:89); kind 2 = log(close / close.shift(1)), +-inf -> 0, row 0 NaN
 *   (:90-91).  d_open is read by kind 1 only. */
int gpb_prep_returns(gpb_handle* h, const double* d_close, const double* d_open, int64_t T, int64_t A,
                     int kind, double* d_out);
/* Per-column mean and standard deviation (ddof = 1 is pandas' Series.std, ddof = 0 numpy's) and, if
 * d_out is not NULL, d_out[t*ldo + a] = (x[t,a] - mean[a]) / std[a]: with ldo > A and an offset
 * d_out pointer several z-scored blocks land side by side in one [T, D] design matrix
 * (concatenate_X).  d_mean / d_std [A] may be NULL. */
int gpb_prep_zscore(gpb_handle* h, const double* d_x, int64_t T, int64_t A, int ddof, double* d_out,
                    int64_t ldo, double* d_mean, double* d_std);
/* Window gather: d_feat [S, T, D], d_y [S, T] (or NULL)  ->  d_X [S*W, N, D], d_Y [S*W, N] with
 * W = (T - N) / stride + 1 windows per series, window w covering rows w*stride .. w*stride+N-1:
 * the [B, N, D] batch gpb_batched_lml_grad consumes.  T < N produces no window (returns 0). */
int gpb_prep_windows(gpb_handle* h, const double* d_feat, const double* d_y, int64_t S, int64_t T, int D,
                     int64_t N, int64_t stride, double* d_X, double* d_Y);

/* ---- prediction post-processing (the step right after predict_f) -----------------------
 * Replaces Predictor.upsample_predictions (/root/reference GPR/predictor.py:35-51):
 * pd.Series(pred, index=X).reindex(X_daily).interpolate('linear') for Q prediction columns at once
 * (d_pred [Q, Ns] -> d_out [Q, Nd]): exact-value lookup of the sparse points in the daily grid, linear
 * interpolation over daily POSITIONS, NaN before the first matched point, last value repeated after
 * the last.  d_Xdaily [Nd] and d_X [Ns] ascending and unique.  Bit-identical to pandas. */
int gpb_post_upsample(gpb_handle* h, const double* d_Xdaily, int64_t Nd, const double* d_X, int64_t Ns,
                      const double* d_pred, int Q, double* d_out);
/* Replaces the blend of Predictor.predict_combined (GPR/predictor.py:27-31):
 * d_out = alpha * daily + beta * weekly + (1 - alpha - beta) * monthly over n values (all four columns
 * of predict_single may be passed as one [4 n] array), the reference's evaluation order. */
int gpb_post_blend(gpb_handle* h, double alpha, double beta, const double* d_daily, const double* d_weekly,
                   const double* d_monthly, int64_t n, double* d_out);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H */
