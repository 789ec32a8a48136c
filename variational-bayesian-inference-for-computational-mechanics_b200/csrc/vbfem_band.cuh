// vbfem_band.cuh -- warp-synchronous banded LDL^T building blocks (sm_100a).
//
// A "front" is one warp eliminating the columns of a banded SPD matrix in LOCAL
// coordinates (band[c*P + k] = A[c+k][c], k = 0..B, P = B+1).  Lane l owns the
// rows congruent to l modulo 32.  The active (B+1)x(B+1) window lives in
// registers: each lane keeps its row's band entries in P accumulator slots
// indexed by column mod P, so that with the column loop unrolled P times every
// register index is static.  A lane whose row has just been eliminated idles
// for 32-P steps before its next row (r+32) becomes active; it uses those
// steps to reload its slots from the assembled band (which also discards the
// by-design garbage that rank-1 updates beyond the diagonal left there).  The
// pivot column is exchanged through shared memory: it is stored exactly where
// the factor L lives (overwriting K in place) and read back as broadcast
// 128-bit loads.  There is no block barrier and no bounds check in the column
// loop: the host pads the system so that every front runs whole blocks of P
// columns (vbfem.cu, "twisted layout").
//
// Two fronts run concurrently on one matrix (twisted factorisation): the top
// front on columns [0, pT) and, on the mirrored numbering, the bottom front on
// the last nB columns; their Schur complements meet in the nm = P middle rows,
// which the top front then finishes.  The triangular sweeps process four rows
// per step (the 4x4 diagonal block is solved redundantly by every lane), which
// shortens the dependency chain per row from shuffle+FMA to about a quarter.
// tests/warp_emulator.py is a lane-level NumPy model of the index logic.
//
// Replaces tf.linalg.solve (src/fem_solver_tf.py:137 upstream) and its
// gradient (the adjoint solve reuses the factor).
#pragma once
#include <cuda_runtime.h>

#include "vbfem_math.cuh"

namespace vbfem {

constexpr unsigned kFull = 0xffffffffu;

template <int B>
struct FrontState {
    static constexpr int P = B + 1;                   // accumulator slots = band column stride
    static constexpr int GAP = 32 - P;                // idle steps between two rows of a lane
    static constexpr int NPER = (P + GAP - 1) / GAP;  // slots reloaded per idle step
    static_assert(P % 2 == 0 && GAP >= 1, "half bandwidth must be odd and <= 29");
    double acc[P];
    double zr;
    int k;  // (lane - j) & 31 for the next column j
    // Look-ahead: the rank-1 update of the column eliminated last is applied one step late, in
    // the shadow of the next column's shuffle -> reciprocal -> first-entry chain.  vp is that
    // column's (masked) pivot-column entry of this lane, wp its scaled column entries 2..B.
    double vp;
    double2 wp[P / 2];  // wp[0] unused
    int bad;            // OR of the high words of all pivots (sign bit set = non-positive pivot)
};

// Start of a front at local column 0: lane l takes row l.
template <int B>
__device__ __forceinline__ void front_init(FrontState<B> &st, const double *__restrict__ band,
                                           const double *__restrict__ z, int lane) {
    constexpr int P = B + 1;
#pragma unroll
    for (int s = 0; s < P; ++s) {
        const int t = (lane - s + 2 * P) % P;  // row - column
        const int c = lane - t;
        st.acc[s] = (c >= 0) ? band[c * P + t] : 0.0;
    }
    st.zr = z[lane];
    st.k = lane;
    st.vp = 0.0;
    st.bad = 0;
#pragma unroll
    for (int q = 0; q < P / 2; ++q) st.wp[q] = make_double2(0.0, 0.0);
}

// Predicated shared-memory load straight into an accumulator register (no select / move).
__device__ __forceinline__ void lds_if(double &dst, const double *p, bool pred) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.shared.f64 %0, [%1]; }"
                 : "+d"(dst)
                 : "r"(a), "r"((int)pred)
                 : "memory");
}

// Predicated shared-memory store (a plain `if (p) *q = v;` compiles to a divergent branch).
__device__ __forceinline__ void sts_if(double *p, double v, bool pred) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.f64 [%0], %1; }"
                 :
                 : "r"(a), "d"(v), "r"((int)pred)
                 : "memory");
}

// Apply the pending rank-1 update (column j-1) to the slots of columns j+1 .. j+B-1, where u is
// the slot of column j.
template <int B>
__device__ __forceinline__ void front_apply_pending(FrontState<B> &st, const int U) {
    constexpr int P = B + 1;
#pragma unroll
    for (int q = 1; q < P / 2; ++q) {
        st.acc[(U + 2 * q - 1) % P] = fma(-st.vp, st.wp[q].x, st.acc[(U + 2 * q - 1) % P]);
        st.acc[(U + 2 * q) % P] = fma(-st.vp, st.wp[q].y, st.acc[(U + 2 * q) % P]);
    }
}

// Flush the look-ahead state (call at a block boundary before the window is read or handed over).
template <int B>
__device__ __forceinline__ void front_flush(FrontState<B> &st) {
    front_apply_pending<B>(st, 0);
    st.vp = 0.0;
}

// Eliminate nblk blocks of P local columns starting at jb0 (a multiple of P): L (unit lower,
// sub-diagonals in band[c][1..B]) and 1/d (band[c][0]) overwrite K in place; z[c] receives the
// forward-eliminated right-hand side.  Rows >= nrows do not exist (their lanes carry harmless
// garbage).  Pivot signs are collected in st.bad.
template <int B>
__device__ __forceinline__ void front_eliminate(FrontState<B> &st, double *__restrict__ band, int nrows,
                                                double *__restrict__ z, int jb0, int nblk) {
    constexpr int P = B + 1, NPER = FrontState<B>::NPER;
    int k = st.k;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        const int jb = jb0 + blk * P;
        // The body below is straight-line code (predication only): the warp stays converged between
        // the shuffles, so the pivot column written by all lanes is visible to the broadcast loads
        // that follow it in program order; the asm volatile statements keep their relative order.
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const int j = jb + u;
            double *col = band + j * P;
            const double v = (k <= B) ? st.acc[u] : 0.0;
            const int src = j & 31;
            // critical chain: pivot d and the first sub-diagonal entry v1 -> 1/d -> the next pivot
            // column (slot u+1) is final; everything else runs one column late
            const double d = __shfl_sync(kFull, v, src);
            const double v1 = __shfl_sync(kFull, v, (src + 1) & 31);
            const double zj = __shfl_sync(kFull, st.zr, src);
            front_apply_pending<B>(st, u);  // column j-1, slots u+1 .. u+B-1
            st.bad |= __double2hiint(d);
            const double t1 = v * v1;
            const double rd = fast_rcp3(d);
            st.acc[(u + 1) % P] = fma(-t1, rd, st.acc[(u + 1) % P]);
            const double w = v * rd;
            const double sv = (k == 0) ? rd : w;
            sts_if(col + k, sv, k <= B);
            sts_if(z + j, zj, k == 0);
            {
                const unsigned ca = (unsigned)__cvta_generic_to_shared(col);
#pragma unroll
                for (int q = 1; q < P / 2; ++q)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                 : "=d"(st.wp[q].x), "=d"(st.wp[q].y)
                                 : "r"(ca + 16u * q)
                                 : "memory");
            }
            st.vp = v;
            const double wz = (k >= 1 && k <= B) ? w : 0.0;
            st.zr = fma(-wz, zj, st.zr);
            // idle lanes: reload NPER slots of the row R = j + k that becomes active next.
            // (R - slot) mod P == (k - P) + C(u, t) reduced once, because jb is a multiple of P.
            // Their pending entry is zero (vp = 0 for k > B) and the lane whose row was the pivot
            // row applies its (garbage) pending update before its first reload in program order.
            const int R = j + k;
            const bool ld = (k > B) && (R < nrows);
            const double *rowp = band + R * P;
#pragma unroll
            for (int t = 0; t < NPER; ++t) {
                const int C = ((-(NPER - 1) * u - t) % P + P) % P;
                int tt = (k - P) + C;
                tt -= (tt >= P) ? P : 0;
                lds_if(st.acc[(NPER * u + t) % P], rowp - tt * (P - 1), ld);
            }
            lds_if(st.zr, z + R, ld);
            k = (k - 1) & 31;
        }
    }
    st.k = k;
}

// Bottom front, after eliminating its ncols columns (a multiple of P): write the Schur
// contributions it holds for the P middle rows into its own (zero-initialised) band extension
// columns [ncols, ncols+P) and its right-hand-side contributions into z[ncols + a].
template <int B>
__device__ __forceinline__ void front_dump_middle(const FrontState<B> &st, double *__restrict__ band, int ncols,
                                                  double *__restrict__ z) {
    constexpr int P = B + 1;
    const int k = st.k;
    if (k < P) {
        const int R = ncols + k;
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int t = (k - s + P) % P;
            if (t <= k) band[(R - t) * P + t] = st.acc[s];
        }
        z[R] = st.zr;
    }
}

// Top front at its first middle column pT (a multiple of P): add the bottom front's contributions
// (mirrored: top middle row a <-> bottom local row ncolsB + P-1-a, same band offset).
template <int B>
__device__ __forceinline__ void front_merge_middle(FrontState<B> &st, const double *__restrict__ bandB, int ncolsB,
                                                   const double *__restrict__ zB) {
    constexpr int P = B + 1;
    const int k = st.k;
    if (k < P) {
        const double *src = bandB + (ncolsB + P - 1 - k) * P;
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int t = (k - s + P) % P;
            if (t <= k) st.acc[s] += src[t];
        }
        st.zr += zB[ncolsB + P - 1 - k];
    }
}

// vec[r] *= 1/d_r for rows [0, nrows)
template <int B>
__device__ __forceinline__ void front_scale(const double *__restrict__ band, double *__restrict__ vec, int nrows,
                                            int lane) {
    for (int r = lane; r < nrows; r += 32) vec[r] *= band[r * (B + 1)];
    __syncwarp();
}

// Operands of one 4-row sweep block: the strictly lower part of the 4x4 diagonal block of L
// (uniform addresses) and this lane's four L entries coupling its row to the block.
struct SweepBlk {
    double l10, l20, l21, l30, l31, l32, m0, m1, m2, m3;
};

// In-place forward substitution L z = w on local columns [lo, hi); rows up to nrows receive
// their partial sums (rows >= hi are written back unfinished: they belong to the next stage).
// Four columns per iteration, operands of the next iteration loaded while the current one
// resolves its shuffle -> 3 FMA chain.
template <int B>
__device__ __forceinline__ void front_fwd_sweep(const double *__restrict__ band, double *__restrict__ z, int lo,
                                                int hi, int nrows, int lane) {
    constexpr int P = B + 1;
    if (lo >= hi) return;
    int r = lo + ((lane - lo) & 31);
    double acc = (r < nrows) ? z[r] : 0.0;
    int j = lo;
    const int nblk = (hi - lo) >> 2;
    auto load_blk = [&](int jj, int rr, bool on) {
        SweepBlk s;
        const double *c0 = band + jj * P;
        const int k = (lane - jj) & 31;
        const double *p = c0 + k;
        const bool v = on && rr < nrows;
        s.l10 = on ? c0[1] : 0.0;
        s.l20 = on ? c0[2] : 0.0;
        s.l30 = on ? c0[3] : 0.0;
        s.l21 = on ? c0[P + 1] : 0.0;
        s.l31 = on ? c0[P + 2] : 0.0;
        s.l32 = on ? c0[2 * P + 1] : 0.0;
        s.m0 = (v && k >= 1 && k <= B) ? p[0] : 0.0;
        s.m1 = (v && k >= 2 && k - 1 <= B) ? p[P - 1] : 0.0;
        s.m2 = (v && k >= 3 && k - 2 <= B) ? p[2 * (P - 1)] : 0.0;
        s.m3 = (v && k >= 4 && k - 3 <= B) ? p[3 * (P - 1)] : 0.0;
        return s;
    };
    SweepBlk cur = load_blk(j, r, nblk > 0);
#pragma unroll 1
    for (int b = 0; b < nblk; ++b, j += 4) {
        const double a0 = __shfl_sync(kFull, acc, j & 31), a1 = __shfl_sync(kFull, acc, (j + 1) & 31);
        const double a2 = __shfl_sync(kFull, acc, (j + 2) & 31), a3 = __shfl_sync(kFull, acc, (j + 3) & 31);
        const int k = (lane - j) & 31;
        const bool piv = k < 4;
        const int rn = piv ? r + 32 : r;
        const SweepBlk nxt = load_blk(j + 4, rn, b + 1 < nblk);
        const double fresh = (piv && rn < nrows) ? z[rn] : 0.0;
        const double z0 = a0;
        const double z1 = fma(-cur.l10, z0, a1);
        const double z2 = fma(-cur.l21, z1, fma(-cur.l20, z0, a2));
        const double z3 = fma(-cur.l32, z2, fma(-cur.l31, z1, fma(-cur.l30, z0, a3)));
        acc = fma(-cur.m1, z1, fma(-cur.m0, z0, acc));
        acc = fma(-cur.m3, z3, fma(-cur.m2, z2, acc));
        // branch-free select of this lane's resolved row (a nested ternary here compiles to a
        // divergent jump table)
        double zs = z3;
        zs = (k == 2) ? z2 : zs;
        zs = (k == 1) ? z1 : zs;
        zs = (k == 0) ? z0 : zs;
        if (piv) z[j + k] = zs;
        acc = piv ? fresh : acc;
        r = rn;
        cur = nxt;
    }
#pragma unroll 1
    for (; j < hi; ++j) {
        const int k = (lane - j) & 31;
        double lv = 0.0;
        if (k >= 1 && k <= B && r < nrows) lv = band[j * P + k];
        const double zj = __shfl_sync(kFull, acc, j & 31);
        acc = fma(-lv, zj, acc);
        if (k == 0) {
            z[j] = zj;
            r += 32;
            acc = (r < nrows) ? z[r] : 0.0;
        }
    }
    if (r < nrows) z[r] = acc;  // rows [hi, hi+32): unfinished partial sums
    __syncwarp();
}

// In-place back substitution L^T x = y on local rows hi..lo (descending); rows > hi are final,
// rows < lo receive their partial sums.
template <int B>
__device__ __forceinline__ void front_back_sweep(const double *__restrict__ band, double *__restrict__ x, int hi,
                                                 int lo, int lane) {
    constexpr int P = B + 1;
    if (hi < lo) return;
    int r = hi - ((hi - lane) & 31);
    double acc = (r >= 0) ? x[r] : 0.0;
    int j = hi;
    const int nblk = (hi - lo + 1) >> 2;
    auto load_blk = [&](int jj, int rr, bool on) {
        SweepBlk s;
        // L[jj-p][jj-q] (p < q) = band[(jj-q)*P + (q-p)];  L[jj-q][rr] = band[rr*P + (i-q)]
        const double *c3 = band + (jj - 3) * P;
        const int i = (jj - lane) & 31;
        const double *p = band + rr * P + i;
        const bool v = on && rr >= 0;
        s.l10 = on ? c3[2 * P + 1] : 0.0;
        s.l20 = on ? c3[P + 2] : 0.0;
        s.l21 = on ? c3[P + 1] : 0.0;
        s.l30 = on ? c3[3] : 0.0;
        s.l31 = on ? c3[2] : 0.0;
        s.l32 = on ? c3[1] : 0.0;
        s.m0 = (v && i >= 1 && i <= B) ? p[0] : 0.0;
        s.m1 = (v && i >= 2 && i - 1 <= B) ? p[-1] : 0.0;
        s.m2 = (v && i >= 3 && i - 2 <= B) ? p[-2] : 0.0;
        s.m3 = (v && i >= 4 && i - 3 <= B) ? p[-3] : 0.0;
        return s;
    };
    SweepBlk cur = load_blk(j, r, nblk > 0);
#pragma unroll 1
    for (int b = 0; b < nblk; ++b, j -= 4) {
        const double a0 = __shfl_sync(kFull, acc, j & 31), a1 = __shfl_sync(kFull, acc, (j - 1) & 31);
        const double a2 = __shfl_sync(kFull, acc, (j - 2) & 31), a3 = __shfl_sync(kFull, acc, (j - 3) & 31);
        const int i = (j - lane) & 31;
        const bool piv = i < 4;
        const int rn = piv ? r - 32 : r;
        const SweepBlk nxt = load_blk(j - 4, rn, b + 1 < nblk);
        const double fresh = (piv && rn >= 0) ? x[rn] : 0.0;
        const double x0 = a0;
        const double x1 = fma(-cur.l10, x0, a1);
        const double x2 = fma(-cur.l21, x1, fma(-cur.l20, x0, a2));
        const double x3 = fma(-cur.l32, x2, fma(-cur.l31, x1, fma(-cur.l30, x0, a3)));
        acc = fma(-cur.m1, x1, fma(-cur.m0, x0, acc));
        acc = fma(-cur.m3, x3, fma(-cur.m2, x2, acc));
        double xsel = x3;
        xsel = (i == 2) ? x2 : xsel;
        xsel = (i == 1) ? x1 : xsel;
        xsel = (i == 0) ? x0 : xsel;
        if (piv) x[j - i] = xsel;
        acc = piv ? fresh : acc;
        r = rn;
        cur = nxt;
    }
#pragma unroll 1
    for (; j >= lo; --j) {
        const int i = (j - lane) & 31;
        double lv = 0.0;
        if (i >= 1 && i <= B && r >= 0) lv = band[r * P + i];
        const double xj = __shfl_sync(kFull, acc, j & 31);
        acc = fma(-lv, xj, acc);
        if (i == 0) {
            x[j] = xj;
            r -= 32;
            acc = (r >= 0) ? x[r] : 0.0;
        }
    }
    if (r >= 0) x[r] = acc;  // rows (lo-32, lo): unfinished partial sums
    __syncwarp();
}

// Bottom front: rows [ncols, ncols+P) of x are final (the shared middle, copied from the top
// front); fold them into the partial sums of rows [ncols-B, ncols).
template <int B>
__device__ __forceinline__ void front_apply_known(const double *__restrict__ band, double *__restrict__ x, int ncols,
                                                  int lane) {
    constexpr int P = B + 1;
    const int r = ncols - 1 - lane;
    if (lane < B && r >= 0) {
        double a0 = x[r], a1 = 0.0;
        const double *p = band + r * P;
#pragma unroll
        for (int o = 1; o <= B; o += 2) {
            if (o > lane) a0 = fma(-p[o], x[r + o], a0);
            if (o + 1 <= B && o + 1 > lane) a1 = fma(-p[o + 1], x[r + o + 1], a1);
        }
        x[r] = a0 + a1;
    }
    __syncwarp();
}

}  // namespace vbfem
