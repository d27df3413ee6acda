// block_chol.cuh -- CTA-cooperative dense kernels on ONE matrix resident in shared memory:
// blocked Cholesky, in-place triangular inverse, warp-level DMMA tile products.  Shared by the
// 128x128 leaf of the blocked Cholesky (cholesky.cu) and by the one-GP-per-CTA batched path
// (batched.cu)  (north_star subsystems 2 and 4).  Every routine is called by all threads of the
// CTA; blockDim.x must be a multiple of 32.
//
// Storage convention: row-major, leading dimension SLD = 132 doubles.  132 = 4 (mod 16) makes the
// 64-bit DMMA fragment loads conflict-free for both k-contiguous and m-contiguous operands:
// a half-warp touches addresses (g*132 + q) or (q*132 + g), g,q in 0..3 -> 16 distinct bank pairs.
// Matrix sizes are padded to a multiple of 8 (the DMMA tile edge) with an identity block.
//
// Measured on B200 (tools/leaf_prof.cu): dependent DMMA 26 cycles, fp64 divide 131 cycles, a 64-bit
// shuffle + add 85 cycles.  The pivot chain is therefore kept free of shuffles and divides: warp 0
// factors each 8x8 diagonal block redundantly in registers (rsqrt per pivot), inverts it, and the
// panel solve becomes a DMMA product with that inverse.
#pragma once
#include <cuda_runtime.h>

// phase-timing hooks for tools/leaf_prof.cu (compiled out everywhere else)
#ifndef GPB_POTRF_STAMP
#define GPB_POTRF_DECL
#define GPB_POTRF_STAMP(i)
#endif

#ifndef GPB_INV_LOOP_ROWS
#define GPB_INV_LOOP_ROWS 14   // rows of the inverse built inside the factorisation loop (the rest right after it)
#endif

namespace gpb {

constexpr int SLD = 132;          // shared-memory leading dimension (doubles)
constexpr int TLD = 68;           // leading dimension of the 64 x 64 temporary used by the inverse
constexpr int DLD = 12;           // leading dimension of the stored 8x8 diagonal-block inverses
constexpr int DINV_DOUBLES = 16 * 8 * DLD;   // 16 diagonal blocks of a 128 x 128 matrix

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// linear index t = ti (ti + 1) / 2 + tj  ->  (ti, tj), 0 <= tj <= ti  (small t: float sqrt is exact enough,
// one correction step each way)
__device__ __forceinline__ void tri_tile(int t, int& ti, int& tj) {
    int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    if ((r + 1) * (r + 2) / 2 <= t) ++r;
    if (r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - r * (r + 1) / 2;
}

// acc (8x8 tile, DMMA C-fragment: thread holds C[g][2q], C[g][2q+1]) += sign * A[8 x klen] * B[klen x 8]
//   A(m, k) at A[m * sam + k * sak],  B(k, n) at B[k * sbk + n * sbn];  klen multiple of 4.
__device__ __forceinline__ void warp_tile_mma(double& c0, double& c1, const double* A, int sam, int sak, const double* B,
                                              int sbk, int sbn, int klen, double sign) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* ap = A + g * sam + q * sak;
    const double* bp = B + q * sbk + g * sbn;
#pragma unroll 4
    for (int k = 0; k < klen; k += 4) {
        const double a = sign * ap[k * sak];
        const double b = bp[k * sbk];
        dmma_8x8x4(c0, c1, a, b);
    }
}

// 2x2 register blocking of the same product: c[ri][rj] (four 8x8 tiles of a 16x16 block) +=
// sign * A_ri[8 x (k1-k0)] * B_rj[(k1-k0) x 8], A tile ri at A + ri*8*sam, B tile rj at B + rj*8*sbn.
// Four independent DMMA accumulator chains per warp and half the fragment loads per tile: the
// in-shared-memory routines are bound by DMMA / LDS latency, not by throughput (tools/leaf_prof.cu).
// a0, a1, b0, b1 (warp-uniform) switch tiles off: their fragments are not loaded and their products
// are skipped -- used for ragged edges and for the k-ranges of triangular operands.
__device__ __forceinline__ void warp_mma_2x2(double (&c)[2][2][2], const double* A, int sam, int sak, const double* B,
                                             int sbk, int sbn, int k0, int k1, double sign, bool a0, bool a1, bool b0,
                                             bool b1) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* ap0 = A + g * sam + q * sak;
    const double* ap1 = ap0 + 8 * sam;
    const double* bp0 = B + q * sbk + g * sbn;
    const double* bp1 = bp0 + 8 * sbn;
#pragma unroll 2
    for (int k = k0; k < k1; k += 4) {
        const double x0 = a0 ? sign * ap0[k * sak] : 0.0, x1 = a1 ? sign * ap1[k * sak] : 0.0;
        const double y0 = b0 ? bp0[k * sbk] : 0.0, y1 = b1 ? bp1[k * sbk] : 0.0;
        if (a0 && b0) dmma_8x8x4(c[0][0][0], c[0][0][1], x0, y0);
        if (a0 && b1) dmma_8x8x4(c[0][1][0], c[0][1][1], x0, y1);
        if (a1 && b0) dmma_8x8x4(c[1][0][0], c[1][0][1], x1, y0);
        if (a1 && b1) dmma_8x8x4(c[1][1][0], c[1][1][1], x1, y1);
    }
}

// ---- Cholesky ----------------------------------------------------------------------------------------
// In-place lower Cholesky of S (np x np, np multiple of 8, np <= 128, stride SLD); only the lower
// triangle is read.  Right-looking with 8-wide panels:
//   (a) warp 0 factors the 8x8 diagonal block and inverts the factor, entirely in registers
//       (every lane redundantly: no shuffles on the pivot chain), stores L_pp and M = L_pp^-1;
//   (b) panel  <- panel * M^T        one DMMA tile product per 8 rows;
//   (c) trailing update C -= P P^T   DMMA tiles of the lower triangle; left-looking per tile row (below).
// Warp 0 owns the chain diagonal block -> panel tile 0 -> next diagonal block; the other warps run (b)
// and (c) in a separate loop (block_potrf_lower below).
// On exit the lower triangle holds L, the strict upper triangle of every 8x8 diagonal block is zero
// (the rest of the upper triangle is never written), and dinv (shared, DINV_DOUBLES) holds the
// inverses of the diagonal blocks (block b at dinv + b*8*DLD, row-major, stride DLD, zero upper).
// *fail (shared int) = 1-based index of the first non-positive pivot, 0 if none; a failing pivot is
// replaced by 1 so that no NaNs propagate.

// 1/sqrt(d) for d > 0 in the normal range, branch-free: MUFU.RSQ64H seed (>= 20 bits) and one cubic
// correction y (1 + e/2 + 3 e^2/8), e = 1 - d y^2  (|y^2 d - 1| <= 3e-16 measured, tools/lat_bench.cu;
// 5 dependent FP64 ops, no slow-path call as in the library rsqrt).  d = +inf / NaN give NaN, which the
// next pivot test flags.
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// (a) one warp, every lane redundantly: factor the 8x8 diagonal block at (p, p) in registers, invert the
// factor, store L_pp (zero strict upper) and M = L_pp^-1.
__device__ __forceinline__ void warp_diag_factor(double* S, int p, int* fail, double* M) {
    // One straight-line block (a single warp runs it, so issue latency per instruction is what
    // counts): every lane holds the whole lower triangle in registers (broadcast LDS.128), the pivot
    // chain d_j -> rsqrt -> scale -> d_j+1 is the critical path, and the rank-1 updates plus the rows
    // of the inverse (all entries, static register indices, no selects) fill its latency bubbles.
    const int lane = threadIdx.x & 31;
    double a[8][8], m[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = S[(p + r) * SLD + p + c];
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double d = a[j][j];
        const bool ok = d > 0.0;
        bad |= ok ? 0u : (1u << j);
        d = ok ? d : 1.0;
        const double rs = rsqrt_pos(d);
        a[j][j] = d * rs;
#pragma unroll
        for (int r = j + 1; r < 8; ++r) a[r][j] *= rs;
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
            for (int k = j + 1; k <= r; ++k) a[r][k] = fma(-a[r][j], a[k][j], a[r][k]);
        // row j of the inverse: m_jj = 1/L_jj, m_jc = -m_jj sum_{k=c}^{j-1} L_jk m_kc (newest m last)
        m[j][j] = rs;
#pragma unroll
        for (int c = 0; c < j; ++c) {
            double s = a[j][c] * m[c][c];
#pragma unroll
            for (int k = c + 1; k < j; ++k) s = fma(a[j][k], m[k][c], s);
            m[j][c] = -s * rs;
        }
    }
    if (lane == 0) {
        if (bad != 0u && *fail == 0) *fail = p + __ffs(bad);
        // L_pp with an explicit zero strict upper part (the trailing updates left symmetric garbage
        // there); the inverse block keeps the zeros dinv was initialised with above the diagonal
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double v0 = (2 * h <= r) ? a[r][2 * h] : 0.0, v1 = (2 * h + 1 <= r) ? a[r][2 * h + 1] : 0.0;
                *reinterpret_cast<double2*>(S + (p + r) * SLD + p + 2 * h) = make_double2(v0, v1);
            }
#pragma unroll
            for (int h = 0; 2 * h <= r; ++h) {
                const double v1 = (2 * h + 1 <= r) ? m[r][2 * h + 1] : 0.0;
                *reinterpret_cast<double2*>(M + r * DLD + 2 * h) = make_double2(m[r][2 * h], v1);
            }
        }
    }
}

__device__ __forceinline__ void warp_trailing_tile(double* S, int p, int t) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    int ti, tj;
    tri_tile(t, ti, tj);
    double* C = S + (p + 8 + ti * 8 + g) * SLD + p + 8 + tj * 8 + 2 * q;
    double2 cc = *reinterpret_cast<double2*>(C);
    const double* Pi = S + (p + 8 + ti * 8) * SLD + p;
    const double* Pj = S + (p + 8 + tj * 8) * SLD + p;
    warp_tile_mma(cc.x, cc.y, Pi, SLD, 1, Pj, 1, SLD, 8, -1.0);
    *reinterpret_cast<double2*>(C) = cc;
}

// tile j of row r of W = L^-1 (j < r), given the rows above it and the inverted diagonal blocks:
//   W[r, j] = -M_r ( L[r, j] M_j + sum_{l = j+1}^{r-1} L[r, l] W[l, j] )
// W's off-diagonal tiles live TRANSPOSED in the unused upper triangle of S (W[a][b] at S[b * SLD + a]), its
// diagonal blocks are the M_r in dinv: both operands of the sum are then k-contiguous rows of S (row block r
// below the diagonal, row block j above it), one loop over the columns (j+1) 8 .. r 8.  scr: 8 x DLD doubles
// of per-warp scratch (the product with M_r needs the sum as a B operand).
__device__ __forceinline__ void warp_inv_tile(double* S, const double* dinv, double* scr, int r, int j) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
    const double* ar = S + (r * 8 + g) * SLD + q;
    const double* mj = dinv + j * 8 * DLD;
    dmma_8x8x4(t0, t1, ar[j * 8], mj[q * DLD + g]);
    dmma_8x8x4(u0, u1, ar[j * 8 + 4], mj[(4 + q) * DLD + g]);
    const double* bj = S + (j * 8 + g) * SLD + q;
#pragma unroll 2
    for (int c = (j + 1) * 8; c < r * 8; c += 8) {
        dmma_8x8x4(t0, t1, ar[c], bj[c]);
        dmma_8x8x4(u0, u1, ar[c + 4], bj[c + 4]);
    }
    *reinterpret_cast<double2*>(scr + g * DLD + 2 * q) = make_double2(t0 + u0, t1 + u1);
    __syncwarp();
    const double* mr = dinv + r * 8 * DLD + g * DLD + q;
    double r0 = 0.0, r1 = 0.0;
    dmma_8x8x4(r0, r1, -mr[0], scr[q * DLD + g]);
    dmma_8x8x4(r0, r1, -mr[4], scr[(4 + q) * DLD + g]);
    __syncwarp();
    S[(j * 8 + 2 * q) * SLD + r * 8 + g] = r0;
    S[(j * 8 + 2 * q + 1) * SLD + r * 8 + g] = r1;
}

// INV = false: the factorisation alone.  INV = true (16 warps; T: 16 x 8 x DLD doubles of scratch): the rows of
// W = L^-1 are built BESIDE the factorisation -- row k in step k, by the updating warps whose tile rows are
// already factorised -- instead of by a separate recursive-doubling pass afterwards (14 K cycles at np = 128,
// a third of the factorisation); on exit L is in the lower triangle as before, W as described at warp_inv_tile.
template <bool INV>
__device__ __forceinline__ void block_potrf_core(double* S, int np, int* fail, double* dinv, double* T) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nt8 = np >> 3;
    if (tid == 0) *fail = 0;
    for (int e = tid; e < nt8 * 8 * DLD; e += nt) dinv[e] = 0.0;   // warp_diag_factor writes the lower parts only
    __syncthreads();

    // (b) panel tile <- tile * M^T  (in place: a warp's operand loads complete before its stores)
    auto panel_tile = [&](int p, int ti) {
        const double* M = dinv + (p >> 3) * 8 * DLD;
        double* Pt = S + (p + 8 + ti * 8) * SLD + p;
        double c0 = 0.0, c1 = 0.0;
        warp_tile_mma(c0, c1, Pt, SLD, 1, M, 1, DLD, 8, 1.0);
        __syncwarp();
        *reinterpret_cast<double2*>(Pt + g * SLD + 2 * q) = make_double2(c0, c1);
    };

    if (nwarps == 1) {
        // single warp: plain right-looking loop
        warp_diag_factor(S, 0, fail, dinv);
        __syncwarp();
        for (int p = 0; p + 8 < np; p += 8) {
            const int mt = (np - p - 8) >> 3;
            for (int ti = 0; ti < mt; ++ti) panel_tile(p, ti);
            __syncwarp();
            for (int t = 0; t < mt * (mt + 1) / 2; ++t) warp_trailing_tile(S, p, t);
            __syncwarp();
            warp_diag_factor(S, p + 8, fail, dinv + ((p + 8) >> 3) * 8 * DLD);
            __syncwarp();
        }
        return;
    }

    // Warp-specialised: two separate loops (their register live ranges do not mix), synchronised by named
    // barriers: 1 publishes the panel tiles (warp 0 only ARRIVES, it never waits inside a step), 2 ends a step.
    if (warp == 0) {
        // the critical chain: panel tile 0 -> update of the next diagonal block -> its factorisation
        warp_diag_factor(S, 0, fail, dinv);
        asm volatile("bar.sync 2, %0;" ::"r"(nt) : "memory");
        GPB_POTRF_DECL
        for (int p = 0; p + 8 < np; p += 8) {
            GPB_POTRF_STAMP(0)
            panel_tile(p, 0);
            asm volatile("bar.arrive 1, %0;" ::"r"(nt) : "memory");
            GPB_POTRF_STAMP(1)
            __syncwarp();
            warp_trailing_tile(S, p, 0);
            __syncwarp();
            GPB_POTRF_STAMP(2)
            warp_diag_factor(S, p + 8, fail, dinv + ((p + 8) >> 3) * 8 * DLD);
            GPB_POTRF_STAMP(3)
            asm volatile("bar.sync 2, %0;" ::"r"(nt) : "memory");
            GPB_POTRF_STAMP(4)
        }
    } else {
    // Updating warps.  Warps 4, 8, 12 share warp 0's scheduler and FP64 pipe (DMMA and DFMA issue to the same
    // unit): they only help with the panel products and sit the update out, so that the pivot chain runs
    // uncontended (measured: the in-situ diagonal factor was 35 % slower than the isolated one).
    const int aw = warp - 1 - (warp >> 2), naw = nwarps - ((nwarps + 3) >> 2);   // index among the updating warps
    const bool upd_warp = (warp & 3) != 0;
    // LEFT-LOOKING update (seven or more updating warps): a warp owns one or two tile ROWS (i = 2 + aw, 2 + aw + naw).
    // In step k it (B) applies the newest panel to its tile (i, k+1), which it has carried in registers since the
    // step before, and stores it -- column k+1 is the next panel --, then (C) builds tile (i, k+2) from the stored
    // original and ALL panels 0..k in one product of depth 8 (k + 1) and keeps it for the next step (the diagonal
    // tile (k+2, k+2) goes back to shared memory: warp 0 applies its last update in the look-ahead).
    // Against the right-looking schedule (every tile visited by every panel): the work of a step is one short
    // and one deep product per warp instead of up to twelve depth-8 ones, so it no longer outlasts warp 0's
    // pivot chain in the first half of the factorisation (tools/leaf_prof.cu: warp 0 waited 8-10 K of 43 K
    // cycles at the end-of-step barrier), and a tile is loaded and stored once.
    const bool left_path = (nt8 - 2) <= 2 * naw;
    int rowS[2] = {nt8, nt8};
    double2 accB[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    if (left_path && upd_warp) {
#pragma unroll
        for (int s_ = 0; s_ < 2; ++s_) {
            const int i = 2 + aw + s_ * naw;
            if (i < nt8) {
                rowS[s_] = i;
                accB[s_] = *reinterpret_cast<const double2*>(S + (i * 8 + g) * SLD + 8 + 2 * q);   // tile (i, 1), no panel yet
            }
        }
    }
    asm volatile("bar.sync 2, %0;" ::"r"(nt) : "memory");
    GPB_POTRF_DECL
    for (int p = 0; p + 8 < np; p += 8) {
        const int mt = (np - p - 8) >> 3;
        const int k = p >> 3;   // panel index
        GPB_POTRF_STAMP(0)
        for (int ti = warp; ti < mt; ti += nwarps - 1) panel_tile(p, ti);
        GPB_POTRF_STAMP(1)
        asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
        GPB_POTRF_STAMP(2)
        // (c) trailing update
        if (upd_warp && left_path) {
#pragma unroll
            for (int s_ = 0; s_ < 2; ++s_) {
                const int i = rowS[s_];
                if (i >= nt8 || i < k + 2) continue;        // no such row / row already factorised
                // (B) last update of tile (i, k+1): -= P(i, k) P(k+1, k)^T
                warp_tile_mma(accB[s_].x, accB[s_].y, S + (i * 8) * SLD + p, SLD, 1, S + ((k + 1) * 8) * SLD + p, 1, SLD, 8, -1.0);
                *reinterpret_cast<double2*>(S + (i * 8 + g) * SLD + (k + 1) * 8 + 2 * q) = accB[s_];
                // (C) tile (i, k+2) <- original - sum_{j <= k} P(i, j) P(k+2, j)^T   (two accumulator chains)
                if (k + 2 < nt8) {
                    double2 c = *reinterpret_cast<const double2*>(S + (i * 8 + g) * SLD + (k + 2) * 8 + 2 * q);
                    double d0 = 0.0, d1 = 0.0;
                    const double* ap = S + (i * 8 + g) * SLD + q;
                    const double* bp = S + ((k + 2) * 8 + g) * SLD + q;
#pragma unroll 2
                    for (int kk = 0; kk < 8 * (k + 1); kk += 8) {
                        dmma_8x8x4(c.x, c.y, -ap[kk], bp[kk]);
                        dmma_8x8x4(d0, d1, -ap[kk + 4], bp[kk + 4]);
                    }
                    c.x += d0;
                    c.y += d1;
                    if (i == k + 2) *reinterpret_cast<double2*>(S + (i * 8 + g) * SLD + (k + 2) * 8 + 2 * q) = c;
                    else accB[s_] = c;
                }
            }
        } else if (upd_warp) {
            // fewer updating warps than the left-looking scheme needs: right-looking, C tiles in shared memory; 16x16
            // super-tiles (I >= J) of the trailing tile grid, one per warp and round; tile (0,0) belongs to
            // warp 0, tiles above the diagonal or past the edge are neither loaded nor stored
            const int o = p + 8, st = (mt + 1) >> 1, nst = st * (st + 1) / 2;
            for (int t = aw; t < nst; t += naw) {
                int I, J;
                tri_tile(t, I, J);
                const int ti0 = 2 * I, tj0 = 2 * J;
                const bool a1 = ti0 + 1 < mt, b1 = (J < I) || a1;
                const bool v[2][2] = {{!(I == 0 && J == 0), J < I}, {a1, a1}};
                double c[2][2][2];
#pragma unroll
                for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                    for (int rj = 0; rj < 2; ++rj) {
                        double2 cc = make_double2(0.0, 0.0);
                        if (v[ri][rj]) cc = *reinterpret_cast<const double2*>(S + (o + (ti0 + ri) * 8 + g) * SLD + o + (tj0 + rj) * 8 + 2 * q);
                        c[ri][rj][0] = cc.x;
                        c[ri][rj][1] = cc.y;
                    }
                warp_mma_2x2(c, S + (o + ti0 * 8) * SLD + p, SLD, 1, S + (o + tj0 * 8) * SLD + p, 1, SLD, 0, 8, -1.0, true, a1, true,
                             b1);
#pragma unroll
                for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                    for (int rj = 0; rj < 2; ++rj)
                        if (v[ri][rj])
                            *reinterpret_cast<double2*>(S + (o + (ti0 + ri) * 8 + g) * SLD + o + (tj0 + rj) * 8 + 2 * q) =
                                make_double2(c[ri][rj][0], c[ri][rj][1]);
            }
        }
        if (INV && upd_warp && k >= 1 && k <= GPB_INV_LOOP_ROWS) {
            // row k of the inverse (M_k and the rows above are complete since the last barrier): tile j goes to
            // the updating warp (k - 1 - j) mod naw -- warps whose tile rows are factorised; the deepest sum
            // (j = 0) to the one that retired last
            for (int j = k - 1 - aw; j >= 0; j -= naw) warp_inv_tile(S, dinv, T + warp * 8 * DLD, k, j);
        }
        GPB_POTRF_STAMP(3)
        asm volatile("bar.sync 2, %0;" ::"r"(nt) : "memory");
        GPB_POTRF_STAMP(4)
    }
    }
    if (INV) {
        // the remaining rows of the inverse: nothing left to run beside them, every warp takes tiles
        for (int r = min(nt8 - 1, GPB_INV_LOOP_ROWS + 1); r < nt8; ++r) {
            for (int j = warp; j < r; j += nwarps) warp_inv_tile(S, dinv, T + warp * 8 * DLD, r, j);
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void block_potrf_lower(double* S, int np, int* fail, double* dinv) {
    block_potrf_core<false>(S, np, fail, dinv, nullptr);
}

// Factor and invert (blockDim.x = 512).  On exit: L in the lower triangle of S (diagonal blocks with a zero strict
// upper part), dinv = the inverted diagonal blocks, and the off-diagonal tiles of W = L^-1 transposed above the
// diagonal: W[a][b] = S[b * SLD + a] for a / 8 > b / 8, = dinv[(a / 8) * 8 * DLD + (a % 8) * DLD + b % 8] otherwise.
__device__ __forceinline__ void block_potrf_inv(double* S, int np, int* fail, double* dinv, double* T) {
    block_potrf_core<true>(S, np, fail, dinv, T);
}

// W (as left by block_potrf_inv) -> the lower triangle of S, row-major, overwriting L.  Tile by tile, a warp
// per tile: lanes run along the contiguous direction of the source (two-way bank conflicts at most on either
// side; an element-wise transposing copy has sixteen-way ones).
// sum_i log L_ii over the first n rows from the inverted diagonal blocks (one warp, fixed order; result in lane 0)
__device__ __forceinline__ double warp_logdiag_from_dinv(const double* dinv, int n) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s -= log(dinv[(i >> 3) * 8 * DLD + (i & 7) * (DLD + 1)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    return s;
}

// first_warp: warps below it take no part (the callers' warp 0 sums the log-diagonal meanwhile, from dinv --
// log L_ii = -log M_ii --, which this pass only reads); all threads meet at the closing barrier.
__device__ __forceinline__ void block_w_to_lower(double* S, int np, const double* dinv, int first_warp = 0) {
    const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5) - first_warp, nwarps = (int)(blockDim.x >> 5) - first_warp;
    const int nt8 = np >> 3;
    const int noff = nt8 * (nt8 - 1) / 2;
    for (int t = (warp >= 0 ? warp : noff + nt8); t < noff + nt8; t += nwarps) {
        if (t < noff) {
            int a, j;
            tri_tile(t, a, j);
            const int r = a + 1;                      // tile (r, j), j < r
            const int rr = lane & 7, c0 = lane >> 3;
            const double v0 = S[(j * 8 + c0) * SLD + r * 8 + rr];
            const double v1 = S[(j * 8 + c0 + 4) * SLD + r * 8 + rr];
            S[(r * 8 + rr) * SLD + j * 8 + c0] = v0;
            S[(r * 8 + rr) * SLD + j * 8 + c0 + 4] = v1;
        } else {
            const int r = t - noff, rr = lane >> 2, c2 = (lane & 3) * 2;
            *reinterpret_cast<double2*>(S + (r * 8 + rr) * SLD + r * 8 + c2) =
                *reinterpret_cast<const double2*>(dinv + r * 8 * DLD + rr * DLD + c2);
        }
    }
    __syncthreads();
}

// ---- triangular inverse, in place ------------------------------------------------------------------------
// S (np x np lower triangular) <- S^-1, np a multiple of 8 and np <= 128, given the inverses of its
// 8x8 diagonal blocks in dinv (as left by block_potrf_lower).  Recursive doubling: for b = 8, 16,
// 32, 64 combine neighbouring inverted blocks: W21 = -W22 (L21 W11), both products on DMMA tiles.
// T (shared) is a 64 x TLD scratch.  Blocks that fall outside np are skipped (np need not be a power
// of two: a trailing partial pair has a shorter second block).
__device__ __forceinline__ void block_trtri_lower_inplace(double* S, int np, double* T, const double* dinv) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int g = lane >> 2, q = lane & 3;
    // base: copy the inverted diagonal blocks in (zero strict upper part)
    for (int e = tid; e < (np >> 3) * 64; e += nt) {
        const int blk = e >> 6, r = (e >> 3) & 7, c = e & 7;
        S[(blk * 8 + r) * SLD + blk * 8 + c] = dinv[blk * 8 * DLD + r * DLD + c];
    }
    __syncthreads();
    for (int b = 8; b < np; b <<= 1) {
        const int npairs = (np + 2 * b - 1) / (2 * b);
        const int bt = b >> 3;  // tiles per block edge
        if (bt == 1) {
            // 8x8 blocks: one tile product per pair and phase
            for (int pr = warp; pr < npairs; pr += nwarps) {
                const int r0 = pr * 2 * b;
                if (r0 + b >= np) continue;
                double c0 = 0.0, c1 = 0.0;
                warp_tile_mma(c0, c1, S + (r0 + b) * SLD + r0, SLD, 1, S + r0 * SLD + r0, SLD, 1, b, 1.0);
                *reinterpret_cast<double2*>(T + (pr * b + g) * TLD + 2 * q) = make_double2(c0, c1);
            }
            __syncthreads();
            for (int pr = warp; pr < npairs; pr += nwarps) {
                const int r0 = pr * 2 * b;
                if (r0 + b >= np) continue;
                double c0 = 0.0, c1 = 0.0;
                warp_tile_mma(c0, c1, S + (r0 + b) * SLD + r0 + b, SLD, 1, T + (pr * b) * TLD, TLD, 1, b, -1.0);
                *reinterpret_cast<double2*>(S + (r0 + b + g) * SLD + r0 + 2 * q) = make_double2(c0, c1);
            }
            __syncthreads();
            continue;
        }
        const int hb = bt >> 1, nsup = npairs * hb * hb;   // 16x16 super-tiles per phase
        // T_pr = L21 W11  (W11 lower: column tile tj needs k >= 8 tj only; its upper tiles hold garbage)
        for (int t = warp; t < nsup; t += nwarps) {
            // column super-tile rotated by (row, pair): the k-range of this phase shrinks with J, and warps
            // w, w+4, w+8, w+12 share a scheduler (and its DMMA issue slot) -- without the rotation one
            // scheduler would get all the J = 0 super-tiles (4x the products of the J = 3 ones)
            const int pr = t / (hb * hb), rem = t - pr * hb * hb, I = rem / hb, J = (rem - I * hb + I + pr) % hb;
            const int r0 = pr * 2 * b, ti0 = 2 * I, tj0 = 2 * J;
            if (r0 + b + ti0 * 8 >= np) continue;                 // second block shorter than b (or absent)
            const bool a1 = r0 + b + (ti0 + 1) * 8 < np;
            double c[2][2][2] = {};
            const double* L21 = S + (r0 + b + ti0 * 8) * SLD + r0;
            const double* W11 = S + r0 * SLD + r0 + tj0 * 8;
            warp_mma_2x2(c, L21, SLD, 1, W11, SLD, 1, tj0 * 8, tj0 * 8 + 8, 1.0, true, a1, true, false);
            warp_mma_2x2(c, L21, SLD, 1, W11, SLD, 1, tj0 * 8 + 8, b, 1.0, true, a1, true, true);
#pragma unroll
            for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                for (int rj = 0; rj < 2; ++rj)
                    if (ri == 0 || a1)
                        *reinterpret_cast<double2*>(T + (pr * b + (ti0 + ri) * 8 + g) * TLD + (tj0 + rj) * 8 + 2 * q) =
                            make_double2(c[ri][rj][0], c[ri][rj][1]);
        }
        __syncthreads();
        // W21 = -W22 T  (W22 lower: row tile ti needs k < 8 (ti + 1) only)
        for (int t = warp; t < nsup; t += nwarps) {
            const int pr = t / (hb * hb), rem = t - pr * hb * hb, I = rem / hb, J = rem - I * hb;
            const int r0 = pr * 2 * b, ti0 = 2 * I, tj0 = 2 * J;
            if (r0 + b + ti0 * 8 >= np) continue;
            const bool a1 = r0 + b + (ti0 + 1) * 8 < np;
            double c[2][2][2] = {};
            const double* W22 = S + (r0 + b + ti0 * 8) * SLD + r0 + b;
            const double* Tp = T + (pr * b) * TLD + tj0 * 8;
            warp_mma_2x2(c, W22, SLD, 1, Tp, TLD, 1, 0, (ti0 + 1) * 8, -1.0, true, a1, true, true);
            warp_mma_2x2(c, W22, SLD, 1, Tp, TLD, 1, (ti0 + 1) * 8, (ti0 + 2) * 8, -1.0, false, a1, true, true);
#pragma unroll
            for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                for (int rj = 0; rj < 2; ++rj)
                    if (ri == 0 || a1)
                        *reinterpret_cast<double2*>(S + (r0 + b + (ti0 + ri) * 8 + g) * SLD + r0 + (tj0 + rj) * 8 + 2 * q) =
                            make_double2(c[ri][rj][0], c[ri][rj][1]);
        }
        __syncthreads();
    }
}

}  // namespace gpb
