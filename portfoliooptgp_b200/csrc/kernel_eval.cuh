// kernel_eval.cuh -- device-side evaluation of a composite covariance expression and of its
// hyper-parameter derivatives for ONE pair (x, x').  Shared by the assembly kernels, the fused
// gradient reduction, the batched one-GP-per-CTA kernel and the SVGP kernels, so every path
// evaluates k(x, x') with the same arithmetic.
//
// Arithmetic restated from GPflow 2.9.1 (un-vendored):
//   gpflow/kernels/stationaries.py  SquaredExponential, RationalQuadratic, Matern12/32/52, Exponential
//   gpflow/kernels/linears.py       Linear
//   gpflow/kernels/periodic.py      Periodic
//   gpflow/kernels/base.py          Sum, Product, active_dims slicing
// Deliberate difference: distances use the direct form sum (x_d - x'_d)^2 rather
// than GPflow's Gram form |x|^2 + |x'|^2 - 2 x.x', which is what a fused kernel computes naturally
// and is the more accurate of the two; the oracle reproduces the Gram form to bound the gap.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/gpb200.h"

namespace gpb {

// ---- exp for the kernel evaluations ---------------------------------------------------------------
// exp(x) = 2^k * P(r), k = rint(x log2 e), r = x - k ln2 (two-term Cody-Waite), P = degree-13 Taylor
// polynomial on |r| <= ln2 / 2 (truncation 4e-18, total error about 1 ulp).  The library exp()
// materialises every coefficient with two UMOVs per use and is issue-bound (50 instructions for 15
// FP64 ones); here the coefficients come from the constant bank and a vector of V independent
// arguments shares each coefficient load, which makes the evaluation FP64-pipe-bound.
// Gradual underflow below 2^-1000 is handled by a second scaling; x < -750 returns 0.
static __constant__ double GPB_EXPC[14] = {1.0, 1.0, 0.5, 1.6666666666666666e-01, 4.1666666666666664e-02,
                                    8.333333333333333e-03, 1.388888888888889e-03, 1.984126984126984e-04,
                                    2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
                                    2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10};
static __constant__ double GPB_EXPK[4] = {1.4426950408889634, 6.93147180369123816490e-01, 1.90821492927058770002e-10,
                                   6755399441055744.0};

// FP64 compares / min / max (DSETP, DMNMX) issue on the low-rate XU pipe on sm_100 (ncu: xu 88 % busy with
// them in the element loops), so range checks here are INTEGER tests on the high word of the double.
__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }

template <int V>
__device__ __forceinline__ void exp_vec(double (&x)[V]) {
    double r[V], p[V];
    int k[V];
    const double l2e = GPB_EXPK[0], ln2h = GPB_EXPK[1], ln2l = GPB_EXPK[2], magic = GPB_EXPK[3];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        // |x| >= 704 (incl. inf / NaN) is out of the fast range: whatever this computes for it is
        // discarded by the fix-up branch below (no traps on the device, so garbage in flight is harmless)
        const double t = fma(x[v], l2e, magic);
        k[v] = __double2loint(t);
        const double n = t - magic;
        r[v] = fma(n, -ln2l, fma(n, -ln2h, x[v]));
        p[v] = GPB_EXPC[13];
    }
#pragma unroll
    for (int i = 12; i >= 0; --i) {
        const double c = GPB_EXPC[i];
#pragma unroll
        for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], c);
    }
    int worst = 0;   // one branch for the whole vector instead of one per element
#pragma unroll
    for (int v = 0; v < V; ++v) worst = max(worst, hi_abs(x[v]));
    if (worst >= 0x40860000) {                                          // rare: underflow window, 0, inf, NaN
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const double res = p[v] * __hiloint2double((k[v] + 1023) << 20, 0);
            x[v] = (hi_abs(x[v]) >= 0x40860000) ? exp(x[v]) : res;
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) x[v] = p[v] * __hiloint2double((k[v] + 1023) << 20, 0);   // |k| <= 1016: normal scale
    }
}

__device__ __forceinline__ double gpb_exp_cb(double x) {
    double a[1] = {x};
    exp_vec<1>(a);
    return a[0];
}
__device__ __forceinline__ double gpb_exp(double x) {
    // scalar paths: the library exp (coefficients as immediates) measures a few per cent faster than the
    // constant-bank version here, with the interpreter and with the straight-line shapes alike
#ifdef GPB_SCALAR_EXP_CB
    return gpb_exp_cb(x);
#else
    return exp(x);
#endif
}

struct DevGroup {
    int kind;
    int ard_index;     // theta index of first ARD lengthscale, or -1
    int period_index;  // theta index of period, or -1
    unsigned dim_mask; // active dimensions (bit d); with no ARD lengthscales the weights are exactly this mask
    double inv_period;
    double w[GPB_MAX_DIMS];       // per-dim weight: mask (0/1) or ARD 1/l_d^2 (or 1/l_d for PERIODIC_ABS)
    double inv_ls[GPB_MAX_DIMS];  // ARD only: 1/l_d (for the lengthscale derivative)
    int ard_slot[GPB_MAX_DIMS];   // ARD only: k such that theta[ard_index + k] belongs to dim d
};

struct DevLeaf {
    int kind;
    int group;
    int var_index, ls_index, alpha_index;
    int arg_is_r;     // 1: group value (scaled) is r itself (PERIODIC_ABS); 0: it is r^2 (or the dot product)
    double variance;
    double scale;     // 1/l^2 (r^2 args), 1/l (r args), 1 for ARD / Linear
    double inv_ls;    // 1/l (0 when no scalar lengthscale)
    double alpha;
};

struct DevTerm {
    int n_factors;
    int leaf[GPB_MAX_FACTORS];
};

struct DevKernel {
    int n_dims, n_params, n_groups, n_leaves, n_terms;
    int has_ard;  // any group with per-dimension lengthscales (gradient takes the generic path)
    DevGroup groups[GPB_MAX_GROUPS];
    DevLeaf leaves[GPB_MAX_LEAVES];
    DevTerm terms[GPB_MAX_TERMS];
};

// ---- expression-shape policies --------------------------------------------------------------------
// The evaluation routines below read the STRUCTURE of the expression (how many groups / leaves /
// terms, their kinds, which leaf sits in which term) through a policy class.  DynShape reads it from
// the descriptor at run time (the general interpreter: uniform branches, selects).  StaticShape
// carries it in template arguments, so that after unrolling every kind switch, leaf select and term
// loop folds away and the kernel body is the straight-line arithmetic of that one expression; the
// parameter VALUES (variances, lengthscale weights, active-dimension masks, theta indices) still
// come from the descriptor.  shapes.cuh lists the shapes that are instantiated (the reference's
// kernel list) and matches a descriptor against them.
struct DynShape {
    static constexpr bool is_static = false;
    static constexpr int HAS_KINK = 0;   // unused: the interpreter gets it as a kernel argument
    static constexpr bool FUSED2 = false;
    static constexpr bool UNIT_W = false;
    static __host__ __device__ __forceinline__ int n_groups(const DevKernel& kp) { return kp.n_groups; }
    static __host__ __device__ __forceinline__ int group_kind(const DevKernel& kp, int g) { return kp.groups[g].kind; }
    static __host__ __device__ __forceinline__ int n_leaves(const DevKernel& kp) { return kp.n_leaves; }
    static __host__ __device__ __forceinline__ int leaf_kind(const DevKernel& kp, int l) { return kp.leaves[l].kind; }
    static __host__ __device__ __forceinline__ int leaf_group(const DevKernel& kp, int l) { return kp.leaves[l].group; }
    static __host__ __device__ __forceinline__ int leaf_arg_is_r(const DevKernel& kp, int l) { return kp.leaves[l].arg_is_r; }
    static __host__ __device__ __forceinline__ int n_terms(const DevKernel& kp) { return kp.n_terms; }
    static __host__ __device__ __forceinline__ int term_nf(const DevKernel& kp, int t) { return kp.terms[t].n_factors; }
    static __host__ __device__ __forceinline__ int term_leaf(const DevKernel& kp, int t, int f) { return kp.terms[t].leaf[f]; }
};

// G* = group kind (-1: unused); L* = leaf kind | group << 4 (-1: unused);
// T* = n_factors | leaf0 << 4 | leaf1 << 8 | leaf2 << 12 | leaf3 << 16 (0: unused).  At most 3 groups (the
// third one, G2, is the trailing template argument), 4 leaves (the register-resident gradient path), 4 terms.
template <int G0, int G1, int L0, int L1, int L2, int L3, int T0, int T1, int T2, int T3, int G2 = -1>
struct StaticShape {
    static constexpr bool is_static = true;
    static constexpr bool UNIT_W = false;   // see UnitWeights below
    static constexpr int NG = (G0 >= 0) + (G1 >= 0) + (G2 >= 0);
    static __host__ __device__ constexpr int gcode(int g) { return g == 0 ? G0 : g == 1 ? G1 : G2; }
    static constexpr int NL = (L0 >= 0) + (L1 >= 0) + (L2 >= 0) + (L3 >= 0);
    static constexpr int NT = (T0 != 0) + (T1 != 0) + (T2 != 0) + (T3 != 0);
    static __host__ __device__ constexpr int lcode(int l) { return l == 0 ? L0 : l == 1 ? L1 : l == 2 ? L2 : L3; }
    static __host__ __device__ constexpr int tcode(int t) { return t == 0 ? T0 : t == 1 ? T1 : t == 2 ? T2 : T3; }
    static __host__ __device__ constexpr int kink_of(int c) {   // leaf with a kink at r = 0 (Matern12 / Exponential)
        return (c >= 0 && ((c & 15) == GPB_LEAF_MATERN12 || (c & 15) == GPB_LEAF_EXPONENTIAL)) ? 1 : 0;
    }
    static constexpr int HAS_KINK = kink_of(L0) | kink_of(L1) | kink_of(L2) | kink_of(L3);
    // One product term of two "pure exponential" leaves (SE, Matern12, Exponential) on Euclidean groups:
    // v0 f0 * v1 f1 = v0 v1 exp(arg0 + arg1) needs ONE exponential per element, and every parameter
    // derivative is that value times a rational factor (k1 * k2 of Multi-Input_GPR/main.py:126-135 with
    // both Exponential; SquaredExponential * Matern12 of GPR/main.py:113).
    static __host__ __device__ constexpr bool pure_exp(int c) {
        return c >= 0 && ((c & 15) == GPB_LEAF_SE || (c & 15) == GPB_LEAF_MATERN12 || (c & 15) == GPB_LEAF_EXPONENTIAL);
    }
    static constexpr bool FUSED2 = NL == 2 && NT == 1 && (T0 & 15) == 2 && ((T0 >> 4) & 15) == 0 && ((T0 >> 8) & 15) == 1 &&
                                   pure_exp(L0) && pure_exp(L1) && G0 == GPB_GROUP_EUCLID &&
                                   (G1 < 0 || G1 == GPB_GROUP_EUCLID);
    static __host__ __device__ constexpr int n_groups(const DevKernel&) { return NG; }
    static __host__ __device__ constexpr int group_kind(const DevKernel&, int g) { return gcode(g); }
    static __host__ __device__ constexpr int n_leaves(const DevKernel&) { return NL; }
    static __host__ __device__ constexpr int leaf_kind(const DevKernel&, int l) { return lcode(l) & 15; }
    static __host__ __device__ constexpr int leaf_group(const DevKernel&, int l) { return (lcode(l) >> 4) & 15; }
    static __host__ __device__ constexpr int leaf_arg_is_r(const DevKernel&, int l) {
        return gcode((lcode(l) >> 4) & 15) == GPB_GROUP_PERIODIC_ABS ? 1 : 0;
    }
    static __host__ __device__ constexpr int n_terms(const DevKernel&) { return NT; }
    static __host__ __device__ constexpr int term_nf(const DevKernel&, int t) { return tcode(t) & 15; }
    static __host__ __device__ constexpr int term_leaf(const DevKernel&, int t, int f) { return (tcode(t) >> (4 + 4 * f)) & 15; }
};

// A straight-line shape has no per-dimension lengthscales, so a group's weights are its 0/1 active-dimension
// mask.  Kernels whose descriptor lives in GLOBAL memory (one per GP in the batched kernel) or that are short
// of FP64 issue slots read one mask word and predicate the dimensions instead of loading and multiplying a
// weight per dimension; kernels that get the descriptor as a __grid_constant__ parameter keep the weights
// (there they are free constant-bank operands and the predicates only cost: grad_reduce 0.43 -> 0.46 ms).
template <class SH>
struct UnitWeights : SH {
    static constexpr bool UNIT_W = SH::is_static;
};

// does the descriptor have exactly the structure SH encodes (and no per-dimension lengthscales)?
template <class SH>
inline bool shape_matches(const DevKernel& kp) {
    if (kp.has_ard || kp.n_groups != SH::NG || kp.n_leaves != SH::NL || kp.n_terms != SH::NT) return false;
    for (int g = 0; g < SH::NG; ++g)
        if (kp.groups[g].kind != SH::group_kind(kp, g)) return false;
    for (int l = 0; l < SH::NL; ++l)
        if (kp.leaves[l].kind != SH::leaf_kind(kp, l) || kp.leaves[l].group != SH::leaf_group(kp, l)) return false;
    for (int t = 0; t < SH::NT; ++t) {
        if (kp.terms[t].n_factors != SH::term_nf(kp, t)) return false;
        for (int f = 0; f < kp.terms[t].n_factors; ++f)
            if (kp.terms[t].leaf[f] != SH::term_leaf(kp, t, f)) return false;
    }
    return true;
}

// ---- spec + theta -> DevKernel (host and device: the batched path builds one per GP in shared memory) --
// Returns 0, or 1 + the theta index of a non-positive lengthscale.
__host__ __device__ inline int build_dev_kernel_core(const gpb_kernel_spec& s, const double* theta, DevKernel* out) {
    int bad = 0;
    out->n_dims = s.n_dims; out->n_params = s.n_params;
    out->n_groups = s.n_groups; out->n_leaves = s.n_leaves; out->n_terms = s.n_terms;
    out->has_ard = 0;
    for (int g = 0; g < s.n_groups; ++g) {
        const gpb_group& G = s.groups[g];
        DevGroup& d = out->groups[g];
        d.kind = G.kind; d.ard_index = G.ard_index; d.period_index = G.period_index;
        d.dim_mask = (s.n_dims >= 32) ? G.dim_mask : (G.dim_mask & ((1u << s.n_dims) - 1u));
        d.inv_period = (G.period_index >= 0) ? 1.0 / theta[G.period_index] : 0.0;
        int k = 0;
        for (int dim = 0; dim < GPB_MAX_DIMS; ++dim) {
            d.w[dim] = 0.0; d.inv_ls[dim] = 0.0; d.ard_slot[dim] = 0;
            if (dim < s.n_dims && ((G.dim_mask >> dim) & 1u)) {
                if (G.ard_index >= 0) {
                    out->has_ard = 1;
                    const double l = theta[G.ard_index + k];
                    if (!(l > 0.0) && !bad) bad = 1 + G.ard_index + k;
                    d.w[dim] = (G.kind == GPB_GROUP_PERIODIC_ABS) ? 1.0 / l : 1.0 / (l * l);
                    d.inv_ls[dim] = 1.0 / l;
                    d.ard_slot[dim] = k;
                } else {
                    d.w[dim] = 1.0;
                }
                ++k;
            }
        }
    }
    for (int l = 0; l < s.n_leaves; ++l) {
        const gpb_leaf& L = s.leaves[l];
        DevLeaf& d = out->leaves[l];
        d.kind = L.kind; d.group = L.group;
        d.var_index = L.var_index; d.ls_index = L.ls_index; d.alpha_index = L.alpha_index;
        d.arg_is_r = (s.groups[L.group].kind == GPB_GROUP_PERIODIC_ABS) ? 1 : 0;
        d.variance = theta[L.var_index];
        d.alpha = (L.alpha_index >= 0) ? theta[L.alpha_index] : 1.0;
        if (L.ls_index >= 0) {
            const double ls = theta[L.ls_index];
            if (!(ls > 0.0) && !bad) bad = 1 + L.ls_index;
            d.inv_ls = 1.0 / ls;
            d.scale = d.arg_is_r ? 1.0 / ls : 1.0 / (ls * ls);
        } else {
            d.inv_ls = 0.0;
            d.scale = 1.0;
        }
    }
    for (int t = 0; t < s.n_terms; ++t) {
        out->terms[t].n_factors = s.terms[t].n_factors;
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) out->terms[t].leaf[f] = s.terms[t].leaf[f];
    }
    return bad;
}

// ---- group value -------------------------------------------------------------------------------
// s = reduction over active dims; when GRAD also returns ds/dperiod.
// UNIT: the group has no per-dimension lengthscales (guaranteed for the straight-line shapes), so its weights
// are the 0/1 active-dimension mask: one mask word and predicated dimensions instead of a weight load and a
// multiply per dimension.
template <int DP, bool GRAD, bool UNIT = false>
__device__ __forceinline__ double group_value_k(const DevGroup& g, int kind, const double (&xi)[DP],
                                                const double (&xj)[DP], double& ds_dperiod) {
    double s = 0.0;
    if (GRAD) ds_dperiod = 0.0;
    switch (kind) {
        case GPB_GROUP_EUCLID: {
            if (UNIT) {
                const unsigned mk = g.dim_mask;
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    if ((mk >> d) & 1u) {
                        const double t = xi[d] - xj[d];
                        s = fma(t, t, s);
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    double t = xi[d] - xj[d];
                    s = fma(g.w[d] * t, t, s);
                }
            }
        } break;
        case GPB_GROUP_DOT: {
            if (UNIT) {
                const unsigned mk = g.dim_mask;
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if ((mk >> d) & 1u) s = fma(xi[d], xj[d], s);
            } else {
#pragma unroll
                for (int d = 0; d < DP; ++d) s = fma(g.w[d] * xi[d], xj[d], s);
            }
        } break;
        case GPB_GROUP_PERIODIC_SQ: {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (g.w[d] != 0.0) {
                    double t = xi[d] - xj[d];
                    double sn, cs;
                    sincospi(t * g.inv_period, &sn, &cs);
                    s = fma(g.w[d] * sn, sn, s);
                    // d/dp sin^2(pi t/p) = 2 sin cos * (-pi t / p^2)
                    if (GRAD) ds_dperiod = fma(g.w[d] * sn * cs, -2.0 * M_PI * t * g.inv_period * g.inv_period, ds_dperiod);
                }
            }
        } break;
        default: {  // GPB_GROUP_PERIODIC_ABS
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (g.w[d] != 0.0) {
                    double t = xi[d] - xj[d];
                    double sn, cs;
                    sincospi(t * g.inv_period, &sn, &cs);
                    s = fma(g.w[d], fabs(sn), s);
                    if (GRAD) {
                        double sg = (sn > 0.0) ? 1.0 : ((sn < 0.0) ? -1.0 : 0.0);
                        ds_dperiod = fma(g.w[d] * sg * cs, -M_PI * t * g.inv_period * g.inv_period, ds_dperiod);
                    }
                }
            }
        } break;
    }
    return s;
}

template <int DP, bool GRAD>
__device__ __forceinline__ double group_value(const DevGroup& g, const double (&xi)[DP], const double (&xj)[DP],
                                              double& ds_dperiod) {
    return group_value_k<DP, GRAD>(g, g.kind, xi, xj, ds_dperiod);
}

// ---- leaf value ----------------------------------------------------------------------------------
// v = variance * f(u), u = s * scale.  Returns v; when GRAD also
//   f_out      = f(u)                       (dv/dvariance)
//   dv_du_u    = variance * f'(u) * u       (finite everywhere; lengthscale derivative = that * c / l)
//   dv_ds      = variance * f'(u) * scale   (chain to group parameters; 0 at the r = 0 singularity of
//                                            Matern12/Exponential, matching TF's zero sub-gradient of
//                                            maximum(r2, 1e-36))
//   dv_dalpha  (RationalQuadratic)
struct LeafOut {
    double v, f, dv_du_u, dv_ds, dv_dalpha;
};

template <bool GRAD>
__device__ __forceinline__ LeafOut leaf_value_k(const DevLeaf& lf, const int kind, const int arg_is_r, double s) {
    LeafOut o;
    o.dv_dalpha = 0.0;
    const double u = s * lf.scale;
    double f, fp_u, fp;  // f(u), f'(u)*u, f'(u)
    switch (kind) {
        case GPB_LEAF_LINEAR: {
            f = u; fp_u = u; fp = 1.0;
        } break;
        case GPB_LEAF_SE: {
            f = gpb_exp(-0.5 * u);
            fp = -0.5 * f; fp_u = fp * u;
        } break;
        case GPB_LEAF_RQ: {
            const double b = 1.0 + 0.5 * u / lf.alpha;
            const double lb = log(b);
            f = gpb_exp(-lf.alpha * lb);
            fp = -0.5 * f / b; fp_u = fp * u;
            if (GRAD) o.dv_dalpha = lf.variance * f * (-lb + 0.5 * u / (lf.alpha * b));
        } break;
        default: {
            // Matern family: needs r.  From an r^2 argument: r = sqrt(max(r2, 1e-36)) (GPflow K_r2);
            // from a PERIODIC_ABS group the argument already is r (GPflow calls K_r directly).
            const double r = arg_is_r ? u : sqrt(fmax(u, 1e-36));
            double df_dr;  // f'(r)
            if (kind == GPB_LEAF_MATERN12) {
                f = gpb_exp(-r); df_dr = -f;
            } else if (kind == GPB_LEAF_EXPONENTIAL) {
                f = gpb_exp(-0.5 * r); df_dr = -0.5 * f;
            } else if (kind == GPB_LEAF_MATERN32) {
                const double s3 = 1.7320508075688772;
                const double e = gpb_exp(-s3 * r);
                f = (1.0 + s3 * r) * e; df_dr = -3.0 * r * e;
            } else {  // MATERN52
                const double s5 = 2.23606797749979;
                const double e = gpb_exp(-s5 * r);
                f = (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
                df_dr = -(5.0 / 3.0) * r * (1.0 + s5 * r) * e;
            }
            if (arg_is_r) {
                fp = df_dr; fp_u = df_dr * r;
            } else {
                // d/du = df/dr / (2 r); times u = r^2 -> df/dr * r / 2
                fp_u = 0.5 * df_dr * r;
                fp = (u > 1e-36) ? 0.5 * df_dr / r : 0.0;
                if (!(u > 1e-36)) fp_u = 0.0;
            }
        } break;
    }
    o.v = lf.variance * f;
    if (GRAD) {
        o.f = f;
        o.dv_du_u = lf.variance * fp_u;
        o.dv_ds = lf.variance * fp * lf.scale;
    }
    return o;
}

template <bool GRAD>
__device__ __forceinline__ LeafOut leaf_value(const DevLeaf& lf, double s) {
    return leaf_value_k<GRAD>(lf, lf.kind, lf.arg_is_r, s);
}

// ---- forward only: k(x, x') ----------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ double kernel_value(const DevKernel& kp, const double (&xi)[DP], const double (&xj)[DP]) {
    double total = 0.0;
    int cached_group = -1;
    double cached_s = 0.0;
    for (int t = 0; t < kp.n_terms; ++t) {
        const DevTerm& tm = kp.terms[t];
        double prod = 1.0;
        for (int f = 0; f < tm.n_factors; ++f) {
            const DevLeaf& lf = kp.leaves[tm.leaf[f]];
            if (lf.group != cached_group) {
                double dummy;
                cached_s = group_value<DP, false>(kp.groups[lf.group], xi, xj, dummy);
                cached_group = lf.group;
            }
            prod *= leaf_value<false>(lf, cached_s).v;
        }
        total += prod;
    }
    return total;
}

// ---- value + weighted parameter derivatives --------------------------------------------------------
// acc[p] += wgt * dk/dtheta_p for every constrained parameter p; returns k.
// acc is a per-thread array (local memory; dynamically indexed).
template <int DP>
__device__ __forceinline__ double kernel_value_grad(const DevKernel& kp, const double (&xi)[DP], const double (&xj)[DP],
                                                    double wgt, double* acc) {
    double total = 0.0;
    for (int t = 0; t < kp.n_terms; ++t) {
        const DevTerm& tm = kp.terms[t];
        LeafOut lo[GPB_MAX_FACTORS];
        double dsdp[GPB_MAX_FACTORS];
        double prod = 1.0;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            if (f < tm.n_factors) {
                const DevLeaf& lf = kp.leaves[tm.leaf[f]];
                double s = group_value<DP, true>(kp.groups[lf.group], xi, xj, dsdp[f]);
                lo[f] = leaf_value<true>(lf, s);
                prod *= lo[f].v;
            }
        }
        total += prod;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            if (f < tm.n_factors) {
                double adj = wgt;
#pragma unroll
                for (int f2 = 0; f2 < GPB_MAX_FACTORS; ++f2)
                    if (f2 != f && f2 < tm.n_factors) adj *= lo[f2].v;
                const DevLeaf& lf = kp.leaves[tm.leaf[f]];
                const DevGroup& g = kp.groups[lf.group];
                acc[lf.var_index] += adj * lo[f].f;
                if (lf.ls_index >= 0) {
                    // u = s / l^2 -> du/dl = -2u/l ; u = s / l -> du/dl = -u/l
                    const double c = lf.arg_is_r ? -1.0 : -2.0;
                    acc[lf.ls_index] += adj * lo[f].dv_du_u * c * lf.inv_ls;
                }
                if (lf.alpha_index >= 0) acc[lf.alpha_index] += adj * lo[f].dv_dalpha;
                if (g.period_index >= 0) acc[g.period_index] += adj * lo[f].dv_ds * dsdp[f];
                if (g.ard_index >= 0) {
                    // ARD: s = sum_d w_d q_d with w_d = 1/l_d^2 (or 1/l_d); ds/dl_d = -c' w_d q_d / l_d
                    const bool per = (g.kind == GPB_GROUP_PERIODIC_SQ || g.kind == GPB_GROUP_PERIODIC_ABS);
                    const double c = (g.kind == GPB_GROUP_PERIODIC_ABS) ? -1.0 : -2.0;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        if (g.w[d] != 0.0) {
                            double tdiff = xi[d] - xj[d];
                            double q;
                            if (per) {
                                double sn = sinpi(tdiff * g.inv_period);
                                q = (g.kind == GPB_GROUP_PERIODIC_SQ) ? sn * sn : fabs(sn);
                            } else {
                                q = tdiff * tdiff;
                            }
                            acc[g.ard_index + g.ard_slot[d]] += adj * lo[f].dv_ds * c * g.w[d] * q * g.inv_ls[d];
                        }
                    }
                }
            }
        }
    }
    return total;
}


// ---- register-resident gradient path ---------------------------------------------------------------
// For expressions with <= GRAD_FAST_LEAVES leaves and no ARD group every parameter derivative is
// accumulated in statically indexed registers (4 slots per leaf: variance, scalar lengthscale, alpha,
// period) instead of a dynamically indexed per-thread array; the slots are scattered to theta
// indices once per thread at the end (grad_flush).  Covers every kernel the reference builds.
constexpr int GRAD_FAST_LEAVES = 4;

struct GradAcc {
    double var[GRAD_FAST_LEAVES], ls[GRAD_FAST_LEAVES], alpha[GRAD_FAST_LEAVES], period[GRAD_FAST_LEAVES];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int l = 0; l < GRAD_FAST_LEAVES; ++l) var[l] = ls[l] = alpha[l] = period[l] = 0.0;
    }
};

__device__ __forceinline__ bool grad_fast_ok(const DevKernel& kp) {
    return kp.n_leaves <= GRAD_FAST_LEAVES && !kp.has_ard;
}

__device__ __forceinline__ double sel4(const double (&a)[GRAD_FAST_LEAVES], int id) {
    double r = a[0];
#pragma unroll
    for (int l = 1; l < GRAD_FAST_LEAVES; ++l) r = (id == l) ? a[l] : r;
    return r;
}

// log f(u) of a pure-exponential leaf on a Euclidean group (u = s * scale = r^2) and, when GRAD,
// d log f / d lengthscale:  SE: -u/2, u/l ;  Matern12: -r, r/l ;  Exponential: -r/2, r/(2 l)
// (r clamped at 1e-18 like GPflow's sqrt(max(r2, 1e-36)); the derivative is then 1e-18/l ~ 0).
template <bool GRAD>
__device__ __forceinline__ double pure_exp_arg(const DevLeaf& lf, const int kind, double s, double& darg_dls) {
    const double u = s * lf.scale;
    if (kind == GPB_LEAF_SE) {
        if (GRAD) darg_dls = u * lf.inv_ls;
        return -0.5 * u;
    }
    const double r = (__double2hiint(u) < 0x38754484) ? 1e-18 : sqrt(u);
    const double c = (kind == GPB_LEAF_MATERN12) ? 1.0 : 0.5;
    if (GRAD) darg_dls = c * r * lf.inv_ls;
    return -c * r;
}

// Forward value with the leaves statically unrolled (same structure as the gradient fast path): the
// per-leaf constants sit at compile-time offsets of the kernel-parameter block, so the compiler
// hoists them out of the element loops instead of chasing term -> leaf -> group indices per element.
template <int DP, class SH = DynShape>
__device__ __forceinline__ double kernel_value_fast(const DevKernel& kp, const double (&xi)[DP], const double (&xj)[DP]) {
    double v[GRAD_FAST_LEAVES];
    double s_prev = 0.0;
    int g_prev = -1;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        v[l] = 0.0;
        if (l < SH::n_leaves(kp)) {
            const DevLeaf& lf = kp.leaves[l];
            const int gi = SH::leaf_group(kp, l);
            if (gi != g_prev) {
                double dummy;
                s_prev = group_value_k<DP, false, SH::UNIT_W>(kp.groups[gi], SH::group_kind(kp, gi), xi, xj, dummy);
                g_prev = gi;
            }
            v[l] = leaf_value_k<false>(lf, SH::leaf_kind(kp, l), SH::leaf_arg_is_r(kp, l), s_prev).v;
        }
    }
    double total = 0.0;
    constexpr int TU = SH::is_static ? 8 : 1;
#pragma unroll TU
    for (int t = 0; t < SH::n_terms(kp); ++t) {
        double prod = sel4(v, SH::term_leaf(kp, t, 0));
#pragma unroll
        for (int f = 1; f < GPB_MAX_FACTORS; ++f)
            if (f < SH::term_nf(kp, t)) prod *= sel4(v, SH::term_leaf(kp, t, f));
        total += prod;
    }
    return total;
}

// Vectorised leaf: v[e] = variance * f(s[e] * scale) for V elements at once (uniform switch outside,
// the exponentials of all V elements share the coefficient loads through exp_vec).
template <int V>
__device__ __forceinline__ void leaf_value_vec_k(const DevLeaf& lf, const int kind, const int arg_is_r, const double (&s)[V],
                                                 double (&v)[V]) {
    double a[V];
    switch (kind) {
        case GPB_LEAF_LINEAR: {
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = lf.variance * (s[e] * lf.scale);
        } break;
        case GPB_LEAF_SE: {
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] = -0.5 * (s[e] * lf.scale);
            exp_vec<V>(a);
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = lf.variance * a[e];
        } break;
        case GPB_LEAF_RQ: {
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = leaf_value_k<false>(lf, kind, arg_is_r, s[e]).v;
        } break;
        default: {
            double r[V];
            const double c = (kind == GPB_LEAF_MATERN12) ? 1.0
                             : (kind == GPB_LEAF_EXPONENTIAL) ? 0.5
                             : (kind == GPB_LEAF_MATERN32) ? 1.7320508075688772 : 2.23606797749979;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const double u = s[e] * lf.scale;
                // sqrt(max(u, 1e-36)) with an integer test on the high word (u >= 0 here): hi(1e-36) = 0x38754484
                r[e] = arg_is_r ? u : ((__double2hiint(u) < 0x38754484) ? 1e-18 : sqrt(u));
                a[e] = -c * r[e];
            }
            exp_vec<V>(a);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                double poly = 1.0;
                if (kind == GPB_LEAF_MATERN32) poly = 1.0 + c * r[e];
                else if (kind == GPB_LEAF_MATERN52) poly = 1.0 + c * r[e] + (5.0 / 3.0) * r[e] * r[e];
                v[e] = lf.variance * (poly * a[e]);
            }
        } break;
    }
}

template <int V>
__device__ __forceinline__ void leaf_value_vec(const DevLeaf& lf, const double (&s)[V], double (&v)[V]) {
    leaf_value_vec_k<V>(lf, lf.kind, lf.arg_is_r, s, v);
}

// k for the 2 x 2 block {xa, xb} x {xj0, xj1}: out = {k(xa,xj0), k(xa,xj1), k(xb,xj0), k(xb,xj1)}.
// Same arithmetic per element as kernel_value_fast; four elements advance together.
template <int DP, class SH = DynShape>
__device__ __forceinline__ void kernel_value_2x2(const DevKernel& kp, const double (&xa)[DP], const double (&xb)[DP],
                                                 const double (&xj0)[DP], const double (&xj1)[DP], double (&out)[4]) {
    if constexpr (SH::FUSED2) {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        int g_prev = -1;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int gi = SH::leaf_group(kp, l);
            if (gi != g_prev) {
                double dummy;
                const DevGroup& g = kp.groups[gi];
                s[0] = group_value_k<DP, false, SH::UNIT_W>(g, GPB_GROUP_EUCLID, xa, xj0, dummy);
                s[1] = group_value_k<DP, false, SH::UNIT_W>(g, GPB_GROUP_EUCLID, xa, xj1, dummy);
                s[2] = group_value_k<DP, false, SH::UNIT_W>(g, GPB_GROUP_EUCLID, xb, xj0, dummy);
                s[3] = group_value_k<DP, false, SH::UNIT_W>(g, GPB_GROUP_EUCLID, xb, xj1, dummy);
                g_prev = gi;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                double dummy;
                a[e] += pure_exp_arg<false>(kp.leaves[l], SH::leaf_kind(kp, l), s[e], dummy);
            }
        }
        exp_vec<4>(a);
        const double vv = kp.leaves[0].variance * kp.leaves[1].variance;
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = vv * a[e];
        return;
    }
    double v[GRAD_FAST_LEAVES][4];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int g_prev = -1;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[l][e] = 0.0;
        if (l < SH::n_leaves(kp)) {
            const DevLeaf& lf = kp.leaves[l];
            const int gi = SH::leaf_group(kp, l);
            if (gi != g_prev) {
                double dummy;
                const DevGroup& g = kp.groups[gi];
                const int gk = SH::group_kind(kp, gi);
                s[0] = group_value_k<DP, false, SH::UNIT_W>(g, gk, xa, xj0, dummy);
                s[1] = group_value_k<DP, false, SH::UNIT_W>(g, gk, xa, xj1, dummy);
                s[2] = group_value_k<DP, false, SH::UNIT_W>(g, gk, xb, xj0, dummy);
                s[3] = group_value_k<DP, false, SH::UNIT_W>(g, gk, xb, xj1, dummy);
                g_prev = gi;
            }
            leaf_value_vec_k<4>(lf, SH::leaf_kind(kp, l), SH::leaf_arg_is_r(kp, l), s, v[l]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = 0.0;
    constexpr int TU = SH::is_static ? 8 : 1;
#pragma unroll TU
    for (int t = 0; t < SH::n_terms(kp); ++t) {
        double prod[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) prod[e] = 1.0;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            if (f < SH::term_nf(kp, t)) {
                const int id = SH::term_leaf(kp, t, f);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    double fv = v[0][e];
#pragma unroll
                    for (int l = 1; l < GRAD_FAST_LEAVES; ++l) fv = (id == l) ? v[l][e] : fv;
                    prod[e] *= fv;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] += prod[e];
    }
}

template <int DP, class SH = DynShape>
__device__ __forceinline__ double kernel_value_grad_fast(const DevKernel& kp, const double (&xi)[DP],
                                                         const double (&xj)[DP], double wgt, GradAcc& A) {
    if constexpr (SH::FUSED2) {
        double arg = 0.0, dl[2], s_prev = 0.0;
        int g_prev = -1;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int gi = SH::leaf_group(kp, l);
            if (gi != g_prev) {
                double dummy;
                s_prev = group_value_k<DP, false, SH::UNIT_W>(kp.groups[gi], GPB_GROUP_EUCLID, xi, xj, dummy);
                g_prev = gi;
            }
            arg += pure_exp_arg<true>(kp.leaves[l], SH::leaf_kind(kp, l), s_prev, dl[l]);
        }
        const double f = gpb_exp(arg);
        const double v0 = kp.leaves[0].variance, v1 = kp.leaves[1].variance;
        const double wf = wgt * f, k = v0 * v1 * f, wk = wgt * k;
        A.var[0] = fma(wf, v1, A.var[0]);
        A.var[1] = fma(wf, v0, A.var[1]);
        A.ls[0] = fma(wk, dl[0], A.ls[0]);
        A.ls[1] = fma(wk, dl[1], A.ls[1]);
        return k;
    }
    double v[GRAD_FAST_LEAVES], fval[GRAD_FAST_LEAVES], dls[GRAD_FAST_LEAVES], dal[GRAD_FAST_LEAVES],
        dper[GRAD_FAST_LEAVES], ladj[GRAD_FAST_LEAVES];
    double s_prev = 0.0, dsp_prev = 0.0;
    int g_prev = -1;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        v[l] = 1.0; fval[l] = dls[l] = dal[l] = dper[l] = 0.0; ladj[l] = 0.0;
        if (l < SH::n_leaves(kp)) {
            const DevLeaf& lf = kp.leaves[l];
            const int gi = SH::leaf_group(kp, l), air = SH::leaf_arg_is_r(kp, l);
            if (gi != g_prev) {
                s_prev = group_value_k<DP, true, SH::UNIT_W>(kp.groups[gi], SH::group_kind(kp, gi), xi, xj, dsp_prev);
                g_prev = gi;
            }
            const LeafOut lo = leaf_value_k<true>(lf, SH::leaf_kind(kp, l), air, s_prev);
            v[l] = lo.v;
            fval[l] = lo.f;
            dls[l] = lo.dv_du_u * (air ? -1.0 : -2.0) * lf.inv_ls;   // inv_ls = 0 without a scalar lengthscale
            dal[l] = lo.dv_dalpha;
            dper[l] = lo.dv_ds * dsp_prev;                            // 0 for non-periodic groups
        }
    }
    double total = 0.0;
    constexpr int TU = SH::is_static ? 8 : 1;
#pragma unroll TU
    for (int t = 0; t < SH::n_terms(kp); ++t) {
        const int nf = SH::term_nf(kp, t);
        if (nf == 1) {   // plain summand (the common case): adjoint of the leaf is the weight
            const int id = SH::term_leaf(kp, t, 0);
            total += sel4(v, id);
#pragma unroll
            for (int l = 0; l < GRAD_FAST_LEAVES; ++l) ladj[l] += (id == l) ? wgt : 0.0;
            continue;
        }
        double fv[GPB_MAX_FACTORS];
        double prod = 1.0;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            fv[f] = (f < nf) ? sel4(v, SH::term_leaf(kp, t, f)) : 1.0;
            prod *= fv[f];
        }
        total += prod;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            if (f < nf) {
                double adj = wgt;
#pragma unroll
                for (int f2 = 0; f2 < GPB_MAX_FACTORS; ++f2)
                    if (f2 != f && f2 < nf) adj *= fv[f2];
                const int id = SH::term_leaf(kp, t, f);
#pragma unroll
                for (int l = 0; l < GRAD_FAST_LEAVES; ++l) ladj[l] += (id == l) ? adj : 0.0;
            }
        }
    }
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        if (SH::is_static) {
            // only the slots this expression has: a leaf beyond n_leaves, a Linear leaf's lengthscale,
            // a non-RQ alpha and a non-periodic period never receive anything
            if (l < SH::n_leaves(kp)) {
                const int lk = SH::leaf_kind(kp, l), gk = SH::group_kind(kp, SH::leaf_group(kp, l));
                A.var[l] = fma(ladj[l], fval[l], A.var[l]);
                if (lk != GPB_LEAF_LINEAR) A.ls[l] = fma(ladj[l], dls[l], A.ls[l]);
                if (lk == GPB_LEAF_RQ) A.alpha[l] = fma(ladj[l], dal[l], A.alpha[l]);
                if (gk == GPB_GROUP_PERIODIC_SQ || gk == GPB_GROUP_PERIODIC_ABS) A.period[l] = fma(ladj[l], dper[l], A.period[l]);
            }
        } else {
            A.var[l] = fma(ladj[l], fval[l], A.var[l]);
            A.ls[l] = fma(ladj[l], dls[l], A.ls[l]);
            A.alpha[l] = fma(ladj[l], dal[l], A.alpha[l]);
            A.period[l] = fma(ladj[l], dper[l], A.period[l]);
        }
    }
    return total;
}

// gx[d] += coef * ds/dx_d for one group (derivative w.r.t. the FIRST argument x)
template <int DP>
__device__ __forceinline__ void add_group_dx_k(const DevGroup& g, const int kind, const double (&xi)[DP],
                                               const double (&xj)[DP], double coef, double (&gx)[DP]) {
    switch (kind) {
        case GPB_GROUP_EUCLID: {
#pragma unroll
            for (int d = 0; d < DP; ++d) gx[d] = fma(2.0 * coef * g.w[d], xi[d] - xj[d], gx[d]);
        } break;
        case GPB_GROUP_DOT: {
#pragma unroll
            for (int d = 0; d < DP; ++d) gx[d] = fma(coef * g.w[d], xj[d], gx[d]);
        } break;
        case GPB_GROUP_PERIODIC_SQ: {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (g.w[d] != 0.0) {
                    double sn, cs;
                    sincospi((xi[d] - xj[d]) * g.inv_period, &sn, &cs);
                    gx[d] = fma(coef * g.w[d] * 2.0 * sn * cs, M_PI * g.inv_period, gx[d]);
                }
            }
        } break;
        default: {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (g.w[d] != 0.0) {
                    double sn, cs;
                    sincospi((xi[d] - xj[d]) * g.inv_period, &sn, &cs);
                    const double sg = (sn > 0.0) ? 1.0 : ((sn < 0.0) ? -1.0 : 0.0);
                    gx[d] = fma(coef * g.w[d] * sg * cs, M_PI * g.inv_period, gx[d]);
                }
            }
        } break;
    }
}

// As kernel_value_grad_fast, and additionally gx[d] += wgt * dk(x, x')/dx_d (first argument): the
// inducing-point gradient of the SVGP path (gpflow trains inducing_variable.Z).
template <int DP, class SH = DynShape>
__device__ __forceinline__ double kernel_value_grad_x_fast(const DevKernel& kp, const double (&xi)[DP],
                                                           const double (&xj)[DP], double wgt, GradAcc& A,
                                                           double (&gx)[DP]) {
    double v[GRAD_FAST_LEAVES], fval[GRAD_FAST_LEAVES], dls[GRAD_FAST_LEAVES], dal[GRAD_FAST_LEAVES],
        dper[GRAD_FAST_LEAVES], dvds[GRAD_FAST_LEAVES], ladj[GRAD_FAST_LEAVES];
    double s_prev = 0.0, dsp_prev = 0.0;
    int g_prev = -1;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        v[l] = 1.0; fval[l] = dls[l] = dal[l] = dper[l] = dvds[l] = 0.0; ladj[l] = 0.0;
        if (l < SH::n_leaves(kp)) {
            const DevLeaf& lf = kp.leaves[l];
            const int gi = SH::leaf_group(kp, l), air = SH::leaf_arg_is_r(kp, l);
            if (gi != g_prev) {
                s_prev = group_value_k<DP, true, SH::UNIT_W>(kp.groups[gi], SH::group_kind(kp, gi), xi, xj, dsp_prev);
                g_prev = gi;
            }
            const LeafOut lo = leaf_value_k<true>(lf, SH::leaf_kind(kp, l), air, s_prev);
            v[l] = lo.v;
            fval[l] = lo.f;
            dls[l] = lo.dv_du_u * (air ? -1.0 : -2.0) * lf.inv_ls;
            dal[l] = lo.dv_dalpha;
            dper[l] = lo.dv_ds * dsp_prev;
            dvds[l] = lo.dv_ds;
        }
    }
    double total = 0.0;
    constexpr int TU = SH::is_static ? 8 : 1;
#pragma unroll TU
    for (int t = 0; t < SH::n_terms(kp); ++t) {
        const int nf = SH::term_nf(kp, t);
        if (nf == 1) {   // plain summand (the common case): adjoint of the leaf is the weight
            const int id = SH::term_leaf(kp, t, 0);
            total += sel4(v, id);
#pragma unroll
            for (int l = 0; l < GRAD_FAST_LEAVES; ++l) ladj[l] += (id == l) ? wgt : 0.0;
            continue;
        }
        double fv[GPB_MAX_FACTORS];
        double prod = 1.0;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            fv[f] = (f < nf) ? sel4(v, SH::term_leaf(kp, t, f)) : 1.0;
            prod *= fv[f];
        }
        total += prod;
#pragma unroll
        for (int f = 0; f < GPB_MAX_FACTORS; ++f) {
            if (f < nf) {
                double adj = wgt;
#pragma unroll
                for (int f2 = 0; f2 < GPB_MAX_FACTORS; ++f2)
                    if (f2 != f && f2 < nf) adj *= fv[f2];
                const int id = SH::term_leaf(kp, t, f);
#pragma unroll
                for (int l = 0; l < GRAD_FAST_LEAVES; ++l) ladj[l] += (id == l) ? adj : 0.0;
            }
        }
    }
    double cacc = 0.0;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        if (SH::is_static && l >= SH::n_leaves(kp)) continue;
        A.var[l] = fma(ladj[l], fval[l], A.var[l]);
        A.ls[l] = fma(ladj[l], dls[l], A.ls[l]);
        A.alpha[l] = fma(ladj[l], dal[l], A.alpha[l]);
        A.period[l] = fma(ladj[l], dper[l], A.period[l]);
        if (l < SH::n_leaves(kp)) {
            cacc = fma(ladj[l], dvds[l], cacc);
            const int gi = SH::leaf_group(kp, l);
            const bool last = (l + 1 >= SH::n_leaves(kp)) || (SH::leaf_group(kp, l + 1 < GRAD_FAST_LEAVES ? l + 1 : l) != gi);
            if (last) {
                add_group_dx_k<DP>(kp.groups[gi], SH::group_kind(kp, gi), xi, xj, cacc, gx);
                cacc = 0.0;
            }
        }
    }
    return total;
}

// Warp-reduce the register accumulators (fixed shuffle tree) and let lane 0 add them into out[0..P)
// (a zero-initialised per-warp row, any memory space) at their theta indices, in a fixed order.
template <class SH = DynShape>
__device__ __forceinline__ void grad_flush(const DevKernel& kp, GradAcc& A, double* out) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
        if (SH::is_static && l >= SH::n_leaves(kp)) continue;
        double a = A.var[l], b = A.ls[l], c = A.alpha[l], d = A.period[l];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
            d += __shfl_down_sync(0xffffffffu, d, o);
        }
        if (lane == 0 && l < SH::n_leaves(kp)) {
            const DevLeaf& lf = kp.leaves[l];
            out[lf.var_index] += a;
            if (lf.ls_index >= 0) out[lf.ls_index] += b;
            if (lf.alpha_index >= 0) out[lf.alpha_index] += c;
            const int pi = kp.groups[lf.group].period_index;
            if (pi >= 0) out[pi] += d;
        }
    }
}

// dispatch: statically unrolled path when the expression has <= GRAD_FAST_LEAVES leaves
template <int DP, class SH = DynShape>
__device__ __forceinline__ double kernel_value_auto(const DevKernel& kp, const double (&xi)[DP], const double (&xj)[DP]) {
    if (SH::is_static) return kernel_value_fast<DP, SH>(kp, xi, xj);
    return (kp.n_leaves <= GRAD_FAST_LEAVES) ? kernel_value_fast<DP>(kp, xi, xj) : kernel_value<DP>(kp, xi, xj);
}

// k(x, x) on the diagonal (gpflow K_diag): stationary -> variance, Linear -> sum w_d x_d^2.
template <int DP>
__device__ __forceinline__ double kernel_diag_value(const DevKernel& kp, const double (&xi)[DP]) {
    return kernel_value<DP>(kp, xi, xi);
}

}  // namespace gpb
