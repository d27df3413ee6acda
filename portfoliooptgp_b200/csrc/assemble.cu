// assemble.cu -- fused kernel-matrix assembly and the fused gradient reduction (north_star
// subsystems 1 and 3).
//
// Replaces, in one pass over the output, what GPflow issues as separate TF ops for
// kernel(X) / kernel(X, X2) / K_diag + GPR._add_noise_cov:
// MatMul + broadcasts for the distances, Exp/Sqrt/Sin per leaf, AddN/Mul for Sum/Product and
// set_diag for the noise.  No intermediate [N,N] (or Periodic's [N,N,D]) tensor exists here:
// the X tiles are staged in shared memory, every leaf is evaluated in registers, and the finished
// tile is written once with 16-byte stores.
//
// HBM roofline: algorithmic bytes = 8*N*N2 (full) or 8*N(N+1)/2 (lower) written + 8*D*(N+N2) read.
#include <stdlib.h>

#include "engine.cuh"
#include "shapes.cuh"

namespace gpb {

static inline int pad_dims(int D) {
    int dp = 1;
    while (dp < D) dp <<= 1;
    return dp;
}

constexpr int TILE = 64;       // output tile edge
constexpr int ROWS_PT = 8;     // rows per thread
constexpr int ASM_THREADS = 256;

__device__ __forceinline__ void tri_tile_index(int64_t b, int& ti, int& tj) {
    // b = ti (ti + 1) / 2 + tj, 0 <= tj <= ti
    int t = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((int64_t)(t + 1) * (t + 2) / 2 <= b) ++t;
    while ((int64_t)t * (t + 1) / 2 > b) --t;
    ti = t;
    tj = (int)(b - (int64_t)t * (t + 1) / 2);
}

// mode 0: full cross K(X, X2); 1: lower tiles of K(X, X) (+diag_add); 2: lower tiles + mirrored copy.
template <int DP, class SH = DynShape>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ X, int64_t N,
                const double* __restrict__ X2, int64_t N2, int D, double* __restrict__ Kout, int64_t ldk, int mode,
                double diag_add, int tiles_n) {
    __shared__ double xs_i[TILE][DP];
    __shared__ double xs_j[TILE][DP];
    __shared__ double stage[TILE * TILE];  // mirror staging (mode 2), rotation-swizzled columns

    int ti, tj;
    if (mode == 0) {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    } else {
        tri_tile_index(blockIdx.x, ti, tj);
    }
    const int64_t row0 = (int64_t)ti * TILE, col0 = (int64_t)tj * TILE;
    const int tid = threadIdx.x;

    // stage the two X tiles (zero-padded to DP columns, rows beyond N zero)
    for (int e = tid; e < TILE * DP; e += ASM_THREADS) {
        int r = e / DP, d = e % DP;
        int64_t gi = row0 + r, gj = col0 + r;
        xs_i[r][d] = (gi < N && d < D) ? X[gi * D + d] : 0.0;
        xs_j[r][d] = (gj < N2 && d < D) ? X2[gj * D + d] : 0.0;
    }
    __syncthreads();

    const int tx = tid & 31, ty = tid >> 5;
    const int c0 = 2 * tx;
    double xj0[DP], xj1[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        xj0[d] = xs_j[c0][d];
        xj1[d] = xs_j[c0 + 1][d];
    }
    const bool vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(Kout) & 15) == 0);
    const bool diag_tile = (mode != 0) && (ti == tj);

    const bool fastk = SH::is_static || (kp.n_leaves <= GRAD_FAST_LEAVES);
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PT; rr += 2) {
        const int r = ty * ROWS_PT + rr;
        double xa[DP], xb[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            xa[d] = xs_i[r][d];
            xb[d] = xs_i[r + 1][d];
        }
        double v[4];   // (r, c0) (r, c0+1) (r+1, c0) (r+1, c0+1)
        if (fastk) {
            kernel_value_2x2<DP, SH>(kp, xa, xb, xj0, xj1, v);
        } else {
            v[0] = kernel_value<DP>(kp, xa, xj0);
            v[1] = kernel_value<DP>(kp, xa, xj1);
            v[2] = kernel_value<DP>(kp, xb, xj0);
            v[3] = kernel_value<DP>(kp, xb, xj1);
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int rw = r + h2;
            const int64_t gi = row0 + rw;
            double v0 = v[2 * h2], v1 = v[2 * h2 + 1];
            if (diag_tile) {
                if (rw == c0) v0 += diag_add;
                if (rw == c0 + 1) v1 += diag_add;
            }
            if (mode == 2) {
                stage[rw * TILE + ((c0 + rw) & (TILE - 1))] = v0;
                stage[rw * TILE + ((c0 + 1 + rw) & (TILE - 1))] = v1;
            }
            if (gi < N) {
                const int64_t gj = col0 + c0;
                double* p = Kout + gi * ldk + gj;
                if (vec_ok && gj + 1 < N2) {
                    *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
                } else {
                    if (gj < N2) p[0] = v0;
                    if (gj + 1 < N2) p[1] = v1;
                }
            }
        }
    }
    if (mode == 2 && ti != tj) {
        __syncthreads();
        // transposed tile: out[col0 + r][row0 + c] = stage[c][r]
#pragma unroll 1
        for (int rr = 0; rr < ROWS_PT; ++rr) {
            const int r = ty * ROWS_PT + rr;
            const int64_t gi = col0 + r;
            if (gi < N) {
                const int64_t gj = row0 + c0;
                double v0 = stage[c0 * TILE + ((r + c0) & (TILE - 1))];
                double v1 = stage[(c0 + 1) * TILE + ((r + c0 + 1) & (TILE - 1))];
                double* p = Kout + gi * ldk + gj;
                if (vec_ok && gj + 1 < N) {
                    *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
                } else {
                    if (gj < N) p[0] = v0;
                    if (gj + 1 < N) p[1] = v1;
                }
            }
        }
    }
}

// ---- Gram-form assembly on the FP64 tensor pipe ---------------------------------------------------------
// For expressions whose groups are all EUCLID / DOT (no Periodic) with at most GRAM_GROUPS distance
// groups: the per-group inner products x.x' of a 64 x 64 tile are 8x8x4 DMMA tiles over the (sqrt(w)-
// scaled) coordinates, r^2 = |x|^2 + |x'|^2 - 2 x.x' costs three FP64 operations per element instead
// of 3 D, and the leaves are evaluated four elements at a time (exp_vec).  This is what makes the
// assembly approach the HBM roofline; the Gram form is also what GPflow evaluates.
// Accuracy guard: an element falls back to the direct difference form when the
// cancellation would be visible -- |x|^2 + |x'|^2 > 64 (non-standardised inputs), or, for the
// non-smooth Matern12 / Exponential leaves, r^2 < 1e-8 (|x|^2 + |x'|^2) (near-coincident points).
constexpr int GRAM_GROUPS = 2;

template <class SH>
__host__ __device__ constexpr int gram_groups_of() {
    if constexpr (SH::is_static) return SH::NG < 1 ? 1 : SH::NG;
    else return GRAM_GROUPS;
}

// resident CTAs per SM to compile for: single-leaf straight-line kernels fit 64 registers
#ifndef GPB_GRAM_BLOCKS_1LEAF
#define GPB_GRAM_BLOCKS_1LEAF 4
#endif
template <class SH>
__host__ __device__ constexpr int gram_min_blocks() {
    if constexpr (SH::is_static) return SH::NL == 1 ? GPB_GRAM_BLOCKS_1LEAF : 3;
    else return 3;
}

template <int DP, int GG = GRAM_GROUPS>
struct GramSmem {
    static constexpr int DPP = (DP % 8 == 0) ? DP + 4 : DP;   // row stride: conflict-free fragment loads
    static constexpr int A = 0;                                // [G][64][DPP] scaled rows
    static constexpr int B = A + GG * TILE * DPP;      // [G][64][DPP] scaled cols
    static constexpr int NA = B + GG * TILE * DPP;     // [G][64]
    static constexpr int NB_ = NA + GG * TILE;         // [G][64]
    static constexpr int STAGE = NB_ + GG * TILE;      // [64][64] mirror staging
    static constexpr int TOTAL = STAGE + TILE * TILE;
};

template <int DP, class SH = DynShape>
__global__ void __launch_bounds__(ASM_THREADS, gram_min_blocks<SH>())
assemble_gram_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ X, int64_t N,
                     const double* __restrict__ X2, int64_t N2, int D, double* __restrict__ Kout, int64_t ldk, int mode,
                     double diag_add, int tiles_n, int has_kink_rt) {
    const int has_kink = SH::is_static ? SH::HAS_KINK : has_kink_rt;
    // distance groups held in shared memory / registers: exactly the shape's for a static shape.  The
    // mirror staging tile exists only in mode 2 (the launch sizes the dynamic shared memory accordingly),
    // so that lower / cross launches fit three CTAs per SM.
    constexpr int GG = gram_groups_of<SH>();
    using L = GramSmem<DP, GG>;
    constexpr int DPP = L::DPP;
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm + L::A;
    double* Bs = gsm + L::B;
    double* na = gsm + L::NA;
    double* nb = gsm + L::NB_;

    int ti, tj;
    if (mode == 0) {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    } else {
        tri_tile_index(blockIdx.x, ti, tj);
    }
    const int64_t row0 = (int64_t)ti * TILE, col0 = (int64_t)tj * TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int G = SH::n_groups(kp);

    // stage sqrt(w)-scaled coordinates of both tile sides for every group
    for (int e = tid; e < GG * TILE * DP; e += ASM_THREADS) {
        const int gg = e / (TILE * DP), rem = e - gg * TILE * DP, r = rem / DP, d = rem - r * DP;
        double a = 0.0, b = 0.0;
        if (gg < G) {
            const double sw = sqrt(kp.groups[gg].w[d]);
            const int64_t gi = row0 + r, gj = col0 + r;
            if (gi < N && d < D) a = sw * X[gi * D + d];
            if (gj < N2 && d < D) b = sw * X2[gj * D + d];
        }
        As[(gg * TILE + r) * DPP + d] = a;
        Bs[(gg * TILE + r) * DPP + d] = b;
    }
    __shared__ int nbmax_hi[1];
    if (tid == 0) nbmax_hi[0] = 0;
    __syncthreads();
    for (int e = tid; e < GG * TILE; e += ASM_THREADS) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const double a = As[e * DPP + d], b = Bs[e * DPP + d];
            sa = fma(a, a, sa);
            sb = fma(b, b, sb);
        }
        na[e] = sa;
        nb[e] = sb;
        atomicMax(nbmax_hi, __double2hiint(sb));   // norms are >= 0: their high words order like the values
    }
    __syncthreads();

    const bool vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(Kout) & 15) == 0);
    const bool diag_tile = (mode != 0) && (ti == tj);
    const int r = warp * 8 + g;               // tile row of this thread
    // A fragments of this warp's 8 rows, all groups, all k-steps (reused for the 8 column tiles)
    double afr[GG][DP / 4];
#pragma unroll
    for (int gg = 0; gg < GG; ++gg)
#pragma unroll
        for (int ks = 0; ks < DP / 4; ++ks) afr[gg][ks] = As[(gg * TILE + r) * DPP + ks * 4 + q];

    // expression shape (uniform): a plain sum of leaves needs no products / selects
    bool pure_sum = true;
    double mult[GRAD_FAST_LEAVES] = {0.0, 0.0, 0.0, 0.0};
    constexpr int TU = SH::is_static ? 8 : 1;
#pragma unroll TU
    for (int t = 0; t < SH::n_terms(kp); ++t) {
        if (SH::term_nf(kp, t) != 1) pure_sum = false;
        const int id = SH::term_leaf(kp, t, 0);
#pragma unroll
        for (int l = 0; l < GRAD_FAST_LEAVES; ++l) mult[l] += (id == l) ? 1.0 : 0.0;
    }
    double* const out_row = Kout + (row0 + r) * ldk + col0;

#pragma unroll
    for (int ct = 0; ct < TILE / 8; ct += 2) {
        // four elements of this thread: (r, c), (r, c + 1), (r, c + 8), (r, c + 9) with c = ct*8 + 2q
        const int c = ct * 8 + 2 * q;
        double s[GG][4];
#pragma unroll
        for (int gg = 0; gg < GG; ++gg) {
            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
            if (gg < G) {
#pragma unroll
                for (int ks = 0; ks < DP / 4; ++ks) {
                    const double b0 = Bs[(gg * TILE + ct * 8 + g) * DPP + ks * 4 + q];
                    const double b1 = Bs[(gg * TILE + ct * 8 + 8 + g) * DPP + ks * 4 + q];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(d0), "+d"(d1) : "d"(afr[gg][ks]), "d"(b0));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(d2), "+d"(d3) : "d"(afr[gg][ks]), "d"(b1));
                }
            }
            const double dots[4] = {d0, d1, d2, d3};
            const bool euclid = (gg < G) && (SH::group_kind(kp, gg) == GPB_GROUP_EUCLID);
            const double nar = na[gg * TILE + r];
            // |x|^2 + |x'|^2 <= |x_r|^2 + max_tile |x'|^2: when even that bound is below the cancellation
            // threshold (standardised inputs: always) and the expression has no kink leaf, no element of
            // this thread needs a guard and the per-element tests are skipped
            const bool unguarded = !has_kink && (__double2hiint(nar + __hiloint2double(nbmax_hi[0] + 1, 0)) <= 0x40500000);
            if (euclid) {
                // r^2 = |x|^2 + |x'|^2 - 2 x.x' ; integer tests on the high words (FP64 compares are slow):
                //   negative -> 0 ; cancellation guard |x|^2+|x'|^2 > 64, or (kink kernels) r^2 < 2^-27 (...)
                int guard = 0;
                double nn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int cc = c + (e & 1) + ((e >> 1) << 3);
                    nn[e] = nar + nb[gg * TILE + cc];
                    const double val = fma(-2.0, dots[e], nn[e]);   // may round to a tiny negative: harmless for
                    s[gg][e] = val;                                  // SE / RQ, clamped inside the sqrt leaves
                    if (!unguarded) {
                        const int hv = __double2hiint(val), hn = __double2hiint(nn[e]);
                        guard |= (hn > 0x40500000) | (has_kink & (hv < hn - (27 << 20)));
                    }
                }
                if (guard) {   // one (rare) branch for the four elements: direct differences where needed
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = c + (e & 1) + ((e >> 1) << 3);
                        const int hv = __double2hiint(s[gg][e]), hn = __double2hiint(nn[e]);
                        if (hn > 0x40500000 || (has_kink && hv < hn - (27 << 20))) {
                            double acc2 = 0.0;
#pragma unroll
                            for (int d = 0; d < DP; ++d) {
                                const double t = As[(gg * TILE + r) * DPP + d] - Bs[(gg * TILE + cc) * DPP + d];
                                acc2 = fma(t, t, acc2);
                            }
                            s[gg][e] = acc2;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) s[gg][e] = dots[e];
            }
        }
        // leaves, four elements at a time
        double out[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (SH::FUSED2) {
            // v0 f0 * v1 f1 = v0 v1 exp(arg0 + arg1): one exponential for the product of two leaves
            double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int l = 0; l < 2; ++l) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    double dummy;
                    a[e] += pure_exp_arg<false>(kp.leaves[l], SH::leaf_kind(kp, l),
                                                (GG == 1 || SH::leaf_group(kp, l) == 0) ? s[0][e] : s[GG - 1][e], dummy);
                }
            }
            exp_vec<4>(a);
            const double vv = kp.leaves[0].variance * kp.leaves[1].variance;
#pragma unroll
            for (int e = 0; e < 4; ++e) out[e] = vv * a[e];
        } else if (pure_sum) {
            // K = sum_l mult_l * leaf_l : accumulate straight into the outputs, no per-leaf arrays, no selects
#pragma unroll
            for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
                if (l < SH::n_leaves(kp)) {
                    const DevLeaf& lf = kp.leaves[l];
                    const int lk = SH::leaf_kind(kp, l), air = SH::leaf_arg_is_r(kp, l);
                    double vl[4];
                    if (GG == 1 || SH::leaf_group(kp, l) == 0) leaf_value_vec_k<4>(lf, lk, air, s[0], vl);
                    else leaf_value_vec_k<4>(lf, lk, air, s[GG - 1], vl);
#pragma unroll
                    for (int e = 0; e < 4; ++e) out[e] = fma(mult[l], vl[e], out[e]);
                }
            }
        } else {
            double v[GRAD_FAST_LEAVES][4];
#pragma unroll
            for (int l = 0; l < GRAD_FAST_LEAVES; ++l) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[l][e] = 0.0;
                if (l < SH::n_leaves(kp)) {
                    const DevLeaf& lf = kp.leaves[l];
                    const int lk = SH::leaf_kind(kp, l), air = SH::leaf_arg_is_r(kp, l);
                    if (GG == 1 || SH::leaf_group(kp, l) == 0) leaf_value_vec_k<4>(lf, lk, air, s[0], v[l]);
                    else leaf_value_vec_k<4>(lf, lk, air, s[GG - 1], v[l]);
                }
            }
#pragma unroll TU
            for (int t = 0; t < SH::n_terms(kp); ++t) {
                double prod[4] = {1.0, 1.0, 1.0, 1.0};
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
        const int64_t gi = row0 + r;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int cc = c + 8 * h2;
            double v0 = out[2 * h2], v1 = out[2 * h2 + 1];
            if (diag_tile) {
                if (r == cc) v0 += diag_add;
                if (r == cc + 1) v1 += diag_add;
            }
            if (mode == 2 && ti != tj && gi < N) {
                // mirrored copy straight from the registers: for a fixed column the 8 rows of a warp's g-lanes are
                // 8 consecutive doubles of the transposed row, i.e. every 32-byte sector is written whole (no
                // staging tile, no barrier: the staged mirror cost 25 % of the kernel)
                const int64_t gjm = col0 + cc;
                if (gjm < N) Kout[gjm * ldk + gi] = v0;
                if (gjm + 1 < N) Kout[(gjm + 1) * ldk + gi] = v1;
            }
            if (gi < N) {
                const int64_t gj = col0 + cc;
                double* p = out_row + cc;
                if (vec_ok && gj + 1 < N2) {
                    *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
                } else {
                    if (gj < N2) p[0] = v0;
                    if (gj + 1 < N2) p[1] = v1;
                }
            }
        }
    }
}

template <int DP>
__global__ void kdiag_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ X, int64_t N, int D,
                             double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) xi[d] = (d < D) ? X[i * D + d] : 0.0;
    out[i] = kernel_diag_value<DP>(kp, xi);
}

// ---- fused gradient reduction ----------------------------------------------------------------------
// partial[b][p] = sum over the tile of c_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta_p, p < P;
// partial[b][P] = sum over diagonal elements of (alpha_i^2 - Kinv_ii).   dK/dtheta is recomputed from
// the X tiles and never written.  HBM bytes: one read of the lower triangle of
// Kinv (8 N(N+1)/2).
template <int DP, bool FAST, class SH = DynShape>
__global__ void __launch_bounds__(ASM_THREADS, FAST ? 2 : 1)
grad_reduce_kernel(const __grid_constant__ DevKernel kp, const double* __restrict__ X, int64_t N, int D,
                   const double* __restrict__ Kinv, int64_t ldk, const double* __restrict__ alpha,
                   double* __restrict__ partial) {
    __shared__ double xs_i[TILE][DP];
    __shared__ double xs_j[TILE][DP];
    __shared__ double al_i[TILE], al_j[TILE];
    __shared__ double red[ASM_THREADS / 32][GPB_MAX_PARAMS + 1];

    int ti, tj;
    tri_tile_index(blockIdx.x, ti, tj);
    const int64_t row0 = (int64_t)ti * TILE, col0 = (int64_t)tj * TILE;
    const int tid = threadIdx.x;
    for (int e = tid; e < TILE * DP; e += ASM_THREADS) {
        int r = e / DP, d = e % DP;
        int64_t gi = row0 + r, gj = col0 + r;
        xs_i[r][d] = (gi < N && d < D) ? X[gi * D + d] : 0.0;
        xs_j[r][d] = (gj < N && d < D) ? X[gj * D + d] : 0.0;
    }
    if (tid < TILE) {
        al_i[tid] = (row0 + tid < N) ? alpha[row0 + tid] : 0.0;
        al_j[tid] = (col0 + tid < N) ? alpha[col0 + tid] : 0.0;
    }
    __syncthreads();

    const int P = kp.n_params;
    constexpr bool fast = FAST;   // register accumulators (every reference kernel) vs generic path: two kernels
    GradAcc A;
    A.zero();
    double tr = 0.0;
    double acc[FAST ? 1 : GPB_MAX_PARAMS + 1];
    if (!fast)
        for (int p = 0; p <= P; ++p) acc[p] = 0.0;

    const int tx = tid & 31, ty = tid >> 5;
    const int c0 = 2 * tx;
    double xj0[DP], xj1[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        xj0[d] = xs_j[c0][d];
        xj1[d] = xs_j[c0 + 1][d];
    }
    const bool vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(Kinv) & 15) == 0);
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PT; ++rr) {
        const int r = ty * ROWS_PT + rr;
        const int64_t gi = row0 + r;
        if (gi >= N) continue;
        const int64_t gj = col0 + c0;
        if (gj > gi) continue;  // strictly above the diagonal: both columns skipped
        double k0 = 0.0, k1 = 0.0;
        const double* p = Kinv + gi * ldk + gj;
        if (vec_ok && gj + 1 <= gi) {
            double2 t = *reinterpret_cast<const double2*>(p);
            k0 = t.x; k1 = t.y;
        } else {
            k0 = p[0];
            if (gj + 1 <= gi) k1 = p[1];
        }
        double xi[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) xi[d] = xs_i[r][d];
        const double ai = al_i[r];
        {
            double w = ai * al_j[c0] - k0;
            if (gj == gi) { tr += w; } else { w *= 2.0; }
            if (fast) kernel_value_grad_fast<DP, SH>(kp, xi, xj0, w, A);
            else kernel_value_grad<DP>(kp, xi, xj0, w, acc);
        }
        if (gj + 1 <= gi) {
            double w = ai * al_j[c0 + 1] - k1;
            if (gj + 1 == gi) { tr += w; } else { w *= 2.0; }
            if (fast) kernel_value_grad_fast<DP, SH>(kp, xi, xj1, w, A);
            else kernel_value_grad<DP>(kp, xi, xj1, w, acc);
        }
    }
    // block reduction in a fixed order (deterministic): warp shuffle, then warp 0 sums the 8 rows
    if (fast) {
        for (int p = tx; p <= P; p += 32) red[ty][p] = 0.0;
        __syncwarp();
        grad_flush<SH>(kp, A, red[ty]);
    } else {
        for (int p = 0; p < P; ++p) {
            double v = acc[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (tx == 0) red[ty][p] = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tr += __shfl_down_sync(0xffffffffu, tr, o);
    if (tx == 0) red[ty][P] = tr;
    __syncthreads();
    if (tid <= P) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < ASM_THREADS / 32; ++w) v += red[w][tid];
        partial[(int64_t)blockIdx.x * (GPB_MAX_PARAMS + 1) + tid] = v;
    }
}

// out[p] = sum_b partial[b][p] in a fixed order (one block per p, tree inside the block).
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int64_t nblocks, int stride, double* out) {
    __shared__ double sm[256];
    const int p = blockIdx.x;
    double v = 0.0;
    for (int64_t b = threadIdx.x; b < nblocks; b += 256) v += partial[b * stride + p];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p] = sm[0];
}

#define GPB_DISPATCH_DP(D, CALL)                                   \
    switch (pad_dims(D)) {                                         \
        case 1: { constexpr int DP = 1; CALL; } break;             \
        case 2: { constexpr int DP = 2; CALL; } break;             \
        case 4: { constexpr int DP = 4; CALL; } break;             \
        case 8: { constexpr int DP = 8; CALL; } break;             \
        default: { constexpr int DP = 16; CALL; } break;           \
    }

int launch_assemble(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, const double* d_X2, int64_t N2,
                    int D, double* d_K, int64_t ldk, int mode, double diag_add) {
    if (N <= 0 || N2 <= 0) return 0;
    if (D < 1 || D > GPB_MAX_DIMS) return set_error(h, -2, "assemble: D=%d outside [1,%d]", D, GPB_MAX_DIMS);
    if (mode != 0 && (d_X2 != d_X || N2 != N)) return set_error(h, -2, "assemble: symmetric modes need X2 == X");
    const int tiles_m = (int)((N + TILE - 1) / TILE), tiles_n = (int)((N2 + TILE - 1) / TILE);
    const int64_t nblk = (mode == 0) ? (int64_t)tiles_m * tiles_n : (int64_t)tiles_m * (tiles_m + 1) / 2;
    if (nblk > 0x7fffffffLL) return set_error(h, -2, "assemble: too many tiles");
    ProfScope prof(h, PROF_ASSEMBLE, h->stream);
    // Gram-form / DMMA path: all groups EUCLID or DOT, at most two of them, at most four leaves, D >= 3
    bool gram = (kp.n_groups <= GRAM_GROUPS) && (kp.n_leaves <= GRAD_FAST_LEAVES) && D >= 3 && !getenv("GPB_NO_GRAM");
    int has_kink = 0;
    for (int g = 0; g < kp.n_groups; ++g)
        if (kp.groups[g].kind != GPB_GROUP_EUCLID && kp.groups[g].kind != GPB_GROUP_DOT) gram = false;
    for (int l = 0; l < kp.n_leaves; ++l)
        if (kp.leaves[l].kind == GPB_LEAF_MATERN12 || kp.leaves[l].kind == GPB_LEAF_EXPONENTIAL) has_kink = 1;
    const int shape = h->use_shapes ? match_shape(kp) : SHAPE_NONE;
    if (gram) {
        cudaError_t e = cudaSuccess;
        // one attribute call per instantiation (static flag inside the macro body's scope)
#define GPB_GRAM_LAUNCH(DPV)                                                                                            \
    {                                                                                                                    \
        static bool attr_set[GPB_MAX_DEVICES] = {};                                                                      \
        using SHL = GPB_SH_FOR(DPV);                                                                                     \
        using GS = GramSmem<DPV, gram_groups_of<SHL>()>;                                                                 \
        constexpr int SM = GS::TOTAL * (int)sizeof(double);                                                              \
        const int sm_now = GS::STAGE * (int)sizeof(double);   /* (no mirror staging tile any more) */                    \
        e = ensure_dyn_smem(attr_set, h->device, assemble_gram_kernel<DPV, GPB_SH_FOR(DPV)>, (size_t)SM);                \
        if (e == cudaSuccess)                                                                                            \
            assemble_gram_kernel<DPV, GPB_SH_FOR(DPV)><<<(unsigned)nblk, ASM_THREADS, sm_now, h->stream>>>(              \
                kp, d_X, N, d_X2, N2, D, d_K, ldk, mode, diag_add, tiles_n, has_kink);                                   \
    }
#define GPB_SHAPE_BODY_                                  \
    switch (pad_dims(D)) {                               \
        case 4: GPB_GRAM_LAUNCH(4) break;                \
        case 8: GPB_GRAM_LAUNCH(8) break;                \
        default: GPB_GRAM_LAUNCH(16) break;              \
    }
        GPB_DISPATCH_SHAPE(shape)
#undef GPB_SHAPE_BODY_
#undef GPB_GRAM_LAUNCH
        if (e != cudaSuccess) return check_cuda(h, e, "assemble_gram cudaFuncSetAttribute");
        h->launches += 1;
        return check_cuda(h, cudaGetLastError(), "assemble_gram_kernel launch");
    }
#define GPB_SHAPE_BODY_                                                                                      \
    GPB_DISPATCH_DP(D, (assemble_kernel<DP, GPB_SH_FOR(DP)><<<(unsigned)nblk, ASM_THREADS, 0, h->stream>>>(  \
                           kp, d_X, N, d_X2, N2, D, d_K, ldk, mode, diag_add, tiles_n)));
    GPB_DISPATCH_SHAPE(shape)
#undef GPB_SHAPE_BODY_
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "assemble_kernel launch");
}

int launch_kdiag(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, int D, double* d_out) {
    if (N <= 0) return 0;
    if (D < 1 || D > GPB_MAX_DIMS) return set_error(h, -2, "kdiag: D=%d outside [1,%d]", D, GPB_MAX_DIMS);
    const unsigned blocks = (unsigned)((N + 255) / 256);
    GPB_DISPATCH_DP(D, (kdiag_kernel<DP><<<blocks, 256, 0, h->stream>>>(kp, d_X, N, D, d_out)));
    h->launches += 1;
    return check_cuda(h, cudaGetLastError(), "kdiag_kernel launch");
}

int launch_grad_reduce(gpb_handle* h, const DevKernel& kp, const double* d_X, int64_t N, int D, const double* d_Kinv,
                       int64_t ldk, const double* d_alpha, double* d_out) {
    const int tiles = (int)((N + TILE - 1) / TILE);
    const int64_t nblk = (int64_t)tiles * (tiles + 1) / 2;
    double* partial = workspace(h, BUF_RED, (size_t)nblk * (GPB_MAX_PARAMS + 1) * sizeof(double));
    if (!partial) return -1;
    ProfScope prof(h, PROF_GRAD, h->stream);
    const int shape = h->use_shapes ? match_shape(kp) : SHAPE_NONE;
    if (kp.n_leaves <= GRAD_FAST_LEAVES && !kp.has_ard) {
#define GPB_SHAPE_BODY_                                                                                                     \
    GPB_DISPATCH_DP(D, (grad_reduce_kernel<DP, true, GPB_SH_FOR(DP)><<<(unsigned)nblk, ASM_THREADS, 0, h->stream>>>(kp, d_X, N, D, \
                                                                                                         d_Kinv, ldk,     \
                                                                                                         d_alpha, partial)));
        GPB_DISPATCH_SHAPE(shape)
#undef GPB_SHAPE_BODY_
    } else {
        GPB_DISPATCH_DP(D, (grad_reduce_kernel<DP, false><<<(unsigned)nblk, ASM_THREADS, 0, h->stream>>>(kp, d_X, N, D, d_Kinv,
                                                                                                         ldk, d_alpha, partial)));
    }
    int rc = check_cuda(h, cudaGetLastError(), "grad_reduce_kernel launch");
    if (rc) return rc;
    reduce_partials_kernel<<<kp.n_params + 1, 256, 0, h->stream>>>(partial, nblk, GPB_MAX_PARAMS + 1, d_out);
    h->launches += 2;
    return check_cuda(h, cudaGetLastError(), "reduce_partials_kernel launch");
}

}  // namespace gpb
