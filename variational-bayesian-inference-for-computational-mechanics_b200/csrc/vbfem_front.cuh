// vbfem_front.cuh -- warp-synchronous banded LDL^T building blocks (sm_100a), second generation.
//
// A "front" is one warp eliminating the columns of a banded SPD matrix in LOCAL coordinates
// (band[c*P + t] = A[c+t][c], t = 0..B, P = B+1).  Lane l owns the rows congruent to l modulo 32.
// The active window lives in registers: slot t of a lane always holds column j+t of the lane's
// row, where j is the column being eliminated.  Every rank-1 update writes its result ONE SLOT
// DOWN (cur[t-1] = cur[t] - v*w[t]), so the window slides with the pivot for free: the column
// loop is the same code for every column and is not unrolled (the first generation unrolled it
// P times for static register indices; ncu showed the 44 KB body starving the instruction
// cache).  Other properties, all validated lane by lane in tests/front_emulator.py:
//   * look-ahead: the pivot d and the first sub-diagonal entry are broadcast by two shuffles,
//     1/d and the next pivot column follow immediately; the other B-1 updates of the column run
//     one step late, in the shadow of the next column's shuffle -> reciprocal chain;
//   * the pivot column is exchanged through shared memory exactly where the factor L lives
//     (it overwrites K in place) and is read back as broadcast 128-bit loads;
//   * a lane whose row has been eliminated idles for 32-P steps; it reloads its next row
//     (r+32) through four static slots (5, 12, 19, 26), one entry per slot and step;
//   * no block barrier, no bounds checks, no branches in the column loop;
//   * up to three right-hand sides ride along (forward elimination fused into the
//     factorisation); with NRA = 3 the loop also accumulates z0 . D^-1 z1 and z0 . D^-1 z2, which
//     are rows of A^-1 f when z1, z2 are unit vectors (the observed node needs no back
//     substitution).
// Two fronts run concurrently on one matrix (twisted factorisation): the top front on columns
// [0, pT) and, on the mirrored numbering, the bottom front on the last nB columns; their Schur
// complements meet in the P middle rows, which the top front then finishes.  The triangular
// sweeps process four rows per step for NV vectors at once.
//
// Replaces tf.linalg.solve (src/fem_solver_tf.py:137 upstream) and its gradient.
#pragma once
#include <cuda_runtime.h>

#include "vbfem_math.cuh"

namespace vbfem {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Predicated shared-memory load straight into a register (no select / move; the register keeps
// its value when the predicate is false).
__device__ __forceinline__ void lds_if(double &dst, unsigned addr, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.shared.f64 %0, [%1]; }"
                 : "+d"(dst)
                 : "r"(addr), "r"((int)pred)
                 : "memory");
}
// Predicated shared-memory store (a plain `if (p) *q = v;` compiles to a divergent branch).
__device__ __forceinline__ void sts_if(unsigned addr, double v, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.f64 [%0], %1; }"
                 :
                 : "r"(addr), "d"(v), "r"((int)pred)
                 : "memory");
}

template <int B>
struct FrontState {
    static constexpr int P = B + 1;
    static_assert(B == 25, "the reload schedule (slots 5, 12, 19, 26) is laid out for B = 25");
    double cur[B + 2];   // slot t <-> column j + t of this lane's row
    double zr[3];        // right-hand sides of this lane's row
    double vp;           // late update: this lane's pivot-column entry of the previous column
    double2 wp[P / 2];   // late update: scaled previous pivot column, wp[q] = (w[2q], w[2q+1]); q = 0 unused
    double ydot[2];      // sum_j z0_j z(1+i)_j / d_j
    int k;               // (row - j) & 31 for the next column j
    int R;               // this lane's current row
    unsigned pk;         // shared address of band[j*P + k]
};

template <int B>
__device__ __forceinline__ double wp_at(const FrontState<B> &st, const int i) {
    return (i & 1) ? st.wp[i >> 1].y : st.wp[i >> 1].x;
}

// Start of a front at local column 0: lane l takes row l.  zs = shared address of the first
// right-hand-side vector (local region of this front), vs = byte stride between vectors.
template <int B, int NRA>
__device__ __forceinline__ void front_init(FrontState<B> &st, unsigned band, unsigned zs, unsigned vs, int lane) {
    constexpr int P = B + 1;
#pragma unroll
    for (int t = 0; t < B + 2; ++t) {
        const int e = lane - t;
        st.cur[t] = 0.0;
        lds_if(st.cur[t], band + 8u * (unsigned)(t * P + e), e >= 0 && e <= B);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        st.zr[r] = 0.0;
        if (r < NRA) lds_if(st.zr[r], zs + r * vs + 8u * lane, true);
    }
    st.vp = 0.0;
#pragma unroll
    for (int q = 0; q < P / 2; ++q) st.wp[q] = make_double2(0.0, 0.0);
    st.ydot[0] = st.ydot[1] = 0.0;
    st.k = lane;
    st.R = lane;
    st.pk = band + 8u * lane;
}

// Apply the late update in place (no shift): call before the window is read or handed over.
template <int B>
__device__ __forceinline__ void front_flush(FrontState<B> &st) {
#pragma unroll
    for (int t = 1; t < B; ++t) st.cur[t] = fma(-st.vp, wp_at<B>(st, t + 1), st.cur[t]);
    st.vp = 0.0;
}

// Eliminate local columns [j0, j1): L (unit lower, sub-diagonals in band[c][1..B]) and 1/d
// (band[c][0]) overwrite K in place; z_r[c] receives the forward-eliminated right-hand sides.
// Rows beyond the end of the matrix are not masked: their lanes read whatever lies behind the band
// (the kernel keeps zeros or finite data there) and only ever touch their own registers and the
// unused tail slots of the last columns.  Loads of columns beyond the stored band that belong to
// existing rows must hit zeroed memory.
// The loop is rotated: the broadcasts of column j+1 are issued as soon as its entries are final,
// so their latency overlaps the stores / loads that finish column j.
template <int B, int NRA>
__device__ __forceinline__ void front_eliminate(FrontState<B> &st, unsigned band, unsigned zs, unsigned vs, int j0,
                                                int j1) {
    constexpr int P = B + 1;
    int k = st.k, R = st.R;
    unsigned pk = st.pk;
    unsigned colp = band + 8u * (unsigned)(j0 * P);
    unsigned zp = zs + 8u * (unsigned)j0;
    double v = st.cur[0];
    double d = __shfl_sync(kFull, v, j0 & 31);
    double v1 = __shfl_sync(kFull, v, (j0 + 1) & 31);
    // eliminated right-hand sides travel through their final place in shared memory: the pivot
    // lane stores, everybody reads the broadcast back
    double zj[NRA];
#pragma unroll
    for (int r = 0; r < NRA; ++r) sts_if(zp + r * vs, st.zr[r], k == 0);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NRA; ++r)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(zj[r]) : "r"(zp + r * vs) : "memory");
#pragma unroll 1
    for (int j = j0; j < j1; ++j) {
        const bool sub = (unsigned)(k - 1) < (unsigned)B;  // 1 <= k <= B: a sub-diagonal row of column j
        const double vm = sub ? v : 0.0;
        const double t1 = vm * v1;
        const double rd = fast_rcp3(d);
        // late update of column j-1, written one slot down: the window slides with the pivot.  The
        // first QS slot pairs run in the shadow of the reciprocal chain ...
        constexpr int QS = 4;
#pragma unroll
        for (int t = 1; t <= 2 * QS; ++t) st.cur[t - 1] = fma(-st.vp, wp_at<B>(st, t + 1), st.cur[t]);
        // first entry of column j's update: the next pivot column is final -> broadcast it now
        st.cur[0] = fma(-t1, rd, st.cur[0]);
        const int srcn = (j + 1) & 31;
        const double vn = st.cur[0];
        const double dn = __shfl_sync(kFull, vn, srcn);
        const double v1n = __shfl_sync(kFull, vn, (srcn + 1) & 31);
        // column j: L entries and 1/d to the band
        const double w = vm * rd;
        sts_if(pk, w, sub);
        sts_if(colp, rd, k == 0);
        __syncwarp();  // the scaled column is read back by every lane: order the exchange (CUDA memory model)
        // ... the rest is interleaved with the broadcast loads of column j's scaled entries (each
        // load overwrites a pair the late update has just consumed), so that the FP64 pipe and
        // the shared-memory pipe work at the same time
#pragma unroll
        for (int q = 1; q <= QS; ++q)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                         : "=d"(st.wp[q].x), "=d"(st.wp[q].y)
                         : "r"(colp + 16u * q)
                         : "memory");
        const double vpo = st.vp;
#pragma unroll
        for (int q = QS + 1; q < P / 2; ++q) {
            const double wx = st.wp[q].x, wy = st.wp[q].y;
            st.cur[2 * q - 2] = fma(-vpo, wx, st.cur[2 * q - 1]);
            if (2 * q < B) st.cur[2 * q - 1] = fma(-vpo, wy, st.cur[2 * q]);
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                         : "=d"(st.wp[q].x), "=d"(st.wp[q].y)
                         : "r"(colp + 16u * q)
                         : "memory");
        }
        st.cur[B - 1] = st.cur[B];
        st.cur[B] = st.cur[B + 1];
        st.vp = vm;  // the pivot lane and the lanes between two rows take no part in the late update
#pragma unroll
        for (int r = 0; r < NRA; ++r) st.zr[r] = fma(-w, zj[r], st.zr[r]);
        if (NRA == 3) {
            const double c = zj[0] * rd;
            st.ydot[0] = fma(c, zj[NRA > 1 ? 1 : 0], st.ydot[0]);
            st.ydot[1] = fma(c, zj[NRA > 2 ? 2 : 0], st.ydot[1]);
        }
        // advance to column j+1
        const bool wrap = (k == 0);
        k = (k - 1) & 31;
        R += wrap ? 32 : 0;
        pk += 8u * (P - 1) + (wrap ? 256u : 0u);
        colp += 8u * P;
        zp += 8u;
        // lanes between two rows: one entry per static slot and step (distance e = k - slot)
        const bool pa = (k >= B);
        lds_if(st.cur[B - 20], pk + 8u * (B - 20) * (P - 1), pa && k <= B + 5);
        lds_if(st.cur[B - 13], pk + 8u * (B - 13) * (P - 1), pa);
        lds_if(st.cur[B - 6], pk + 8u * (B - 6) * (P - 1), pa);
        lds_if(st.cur[B + 1], pk + 8u * (B + 1) * (P - 1), k >= B + 1);
#pragma unroll
        for (int r = 0; r < NRA; ++r) lds_if(st.zr[r], zs + r * vs + 8u * (unsigned)R, wrap);
        const bool pst = (k == 0) && (j + 1 < j1);  // nothing to publish behind the last column
#pragma unroll
        for (int r = 0; r < NRA; ++r) sts_if(zp + r * vs, st.zr[r], pst);
        __syncwarp();  // the pivot lane's right-hand-side entries are read by every lane
#pragma unroll
        for (int r = 0; r < NRA; ++r)
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(zj[r]) : "r"(zp + r * vs) : "memory");
        v = vn;
        d = dn;
        v1 = v1n;
    }
    st.k = k;
    st.R = R;
    st.pk = pk;
}

// Bottom front, after eliminating its ncols columns and front_flush: write the Schur
// contributions it holds for the P middle rows (S[k][t], lower triangle, bottom orientation) and
// its three right-hand-side contributions (rs[r][k]).
template <int B>
__device__ __forceinline__ void front_dump_middle(const FrontState<B> &st, unsigned S, unsigned rs) {
    constexpr int P = B + 1;
    const int k = st.k;
#pragma unroll
    for (int t = 0; t < P; ++t) sts_if(S + 8u * (unsigned)(k * P + t), st.cur[t], k < P && t <= k);
#pragma unroll
    for (int r = 0; r < 3; ++r) sts_if(rs + 8u * (unsigned)(r * P + k), st.zr[r], k < P);
}

// Top front at its first middle column (after front_flush): add the bottom front's contributions
// (mirrored: top middle row a, column b <-> bottom S[P-1-b][P-1-a]).
template <int B>
__device__ __forceinline__ void front_merge_middle(FrontState<B> &st, unsigned S, unsigned rs) {
    constexpr int P = B + 1;
    const int k = st.k;
    const bool mid = k < P;
#pragma unroll
    for (int t = 0; t < P; ++t) {
        double a = 0.0;
        lds_if(a, S + 8u * (unsigned)((P - 1 - t) * P + (P - 1 - k)), mid && t <= k);
        st.cur[t] += a;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double a = 0.0;
        lds_if(a, rs + 8u * (unsigned)(r * P + (P - 1 - k)), mid);
        st.zr[r] += a;
    }
}

// x_v[r] *= 1/d_r for rows [lo, hi) of NV vectors (stride vs doubles; NV = 0: check only).  Returns
// non-zero if a pivot of these rows was non-positive or not finite (the elimination loop itself
// does not test them).
template <int B, int NV>
__device__ __forceinline__ int front_scale(const double *__restrict__ band, double *__restrict__ x, int vs,
                                           int lo, int hi, int lane) {
    int bad = 0;
    for (int r = lo + lane; r < hi; r += 32) {
        const double rd = band[r * (B + 1)];
        bad |= !(rd > 0.0 && rd < 1.0e300);
#pragma unroll
        for (int v = 0; v < NV; ++v) x[v * vs + r] *= rd;
    }
    __syncwarp();
    return bad;
}

// Operands of one 4-row sweep block: the strictly lower part of the 4x4 diagonal block of L
// (uniform addresses) and this lane's four L entries coupling its row to the block.
struct SweepBlk {
    double l10, l20, l21, l30, l31, l32, m0, m1, m2, m3;
};

// In-place forward substitution L z = w on local columns [lo, hi) for NV vectors; rows up to
// nrows receive their partial sums (rows >= hi are written back unfinished).
template <int B, int NV>
__device__ __forceinline__ void front_fwd_sweep(const double *__restrict__ band, double *__restrict__ z, int vs,
                                                int lo, int hi, int nrows, int lane) {
    constexpr int P = B + 1;
    if (lo >= hi) return;
    int r = lo + ((lane - lo) & 31);
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = (r < nrows) ? z[v * vs + r] : 0.0;
    int j = lo;
    const int nblk = (hi - lo) >> 2;
    auto load_blk = [&](int jj, int rr, bool on) {
        SweepBlk s;
        const double *c0 = band + jj * P;
        const int k = (lane - jj) & 31;
        const double *p = c0 + k;
        const bool ok = on && rr < nrows;
        s.l10 = on ? c0[1] : 0.0;
        s.l20 = on ? c0[2] : 0.0;
        s.l30 = on ? c0[3] : 0.0;
        s.l21 = on ? c0[P + 1] : 0.0;
        s.l31 = on ? c0[P + 2] : 0.0;
        s.l32 = on ? c0[2 * P + 1] : 0.0;
        s.m0 = (ok && k >= 1 && k <= B) ? p[0] : 0.0;
        s.m1 = (ok && k >= 2 && k - 1 <= B) ? p[P - 1] : 0.0;
        s.m2 = (ok && k >= 3 && k - 2 <= B) ? p[2 * (P - 1)] : 0.0;
        s.m3 = (ok && k >= 4 && k - 3 <= B) ? p[3 * (P - 1)] : 0.0;
        return s;
    };
    auto step = [&](const SweepBlk &cur, SweepBlk &nxt, bool more) {
        const int k = (lane - j) & 31;
        const bool piv = k < 4;
        const int rn = piv ? r + 32 : r;
        nxt = load_blk(j + 4, rn, more);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double a0 = __shfl_sync(kFull, acc[v], j & 31), a1 = __shfl_sync(kFull, acc[v], (j + 1) & 31);
            const double a2 = __shfl_sync(kFull, acc[v], (j + 2) & 31), a3 = __shfl_sync(kFull, acc[v], (j + 3) & 31);
            const double fresh = (piv && rn < nrows) ? z[v * vs + rn] : 0.0;
            const double z0 = a0;
            const double z1 = fma(-cur.l10, z0, a1);
            const double z2 = fma(-cur.l21, z1, fma(-cur.l20, z0, a2));
            const double z3 = fma(-cur.l32, z2, fma(-cur.l31, z1, fma(-cur.l30, z0, a3)));
            double a = fma(-cur.m1, z1, fma(-cur.m0, z0, acc[v]));
            a = fma(-cur.m3, z3, fma(-cur.m2, z2, a));
            double zsel = z3;
            zsel = (k == 2) ? z2 : zsel;
            zsel = (k == 1) ? z1 : zsel;
            zsel = (k == 0) ? z0 : zsel;
            if (piv) z[v * vs + j + k] = zsel;
            acc[v] = piv ? fresh : a;
        }
        r = rn;
        j += 4;
    };
    SweepBlk blkA = load_blk(j, r, nblk > 0), blkB;
    int b = 0;
#pragma unroll 1
    for (; b + 1 < nblk; b += 2) {
        step(blkA, blkB, true);
        step(blkB, blkA, b + 2 < nblk);
    }
    if (b < nblk) step(blkA, blkB, false);
#pragma unroll 1
    for (; j < hi; ++j) {
        const int k = (lane - j) & 31;
        double lv = 0.0;
        if (k >= 1 && k <= B && r < nrows) lv = band[j * P + k];
        const bool piv = (k == 0);
        const int rn = piv ? r + 32 : r;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double zj = __shfl_sync(kFull, acc[v], j & 31);
            const double fresh = (piv && rn < nrows) ? z[v * vs + rn] : 0.0;
            if (piv) z[v * vs + j] = zj;
            acc[v] = piv ? fresh : fma(-lv, zj, acc[v]);
        }
        r = rn;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
        if (r < nrows) z[v * vs + r] = acc[v];  // rows [hi, hi+32): unfinished partial sums
    __syncwarp();
}

// In-place back substitution L^T x = y on local rows hi..lo (descending) for NV vectors; rows
// > hi are final, rows < lo receive their partial sums.
template <int B, int NV>
__device__ __forceinline__ void front_back_sweep(const double *__restrict__ band, double *__restrict__ x, int vs,
                                                 int hi, int lo, int lane) {
    constexpr int P = B + 1;
    if (hi < lo) return;
    int r = hi - ((hi - lane) & 31);
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = (r >= 0) ? x[v * vs + r] : 0.0;
    int j = hi;
    const int nblk = (hi - lo + 1) >> 2;
    auto load_blk = [&](int jj, int rr, bool on) {
        SweepBlk s;
        // L[jj-p][jj-q] (p < q) = band[(jj-q)*P + (q-p)];  L[jj-q][rr] = band[rr*P + (i-q)]
        const double *c3 = band + (jj - 3) * P;
        const int i = (jj - lane) & 31;
        const double *p = band + rr * P + i;
        const bool ok = on && rr >= 0;
        s.l10 = on ? c3[2 * P + 1] : 0.0;
        s.l20 = on ? c3[P + 2] : 0.0;
        s.l21 = on ? c3[P + 1] : 0.0;
        s.l30 = on ? c3[3] : 0.0;
        s.l31 = on ? c3[2] : 0.0;
        s.l32 = on ? c3[1] : 0.0;
        s.m0 = (ok && i >= 1 && i <= B) ? p[0] : 0.0;
        s.m1 = (ok && i >= 2 && i - 1 <= B) ? p[-1] : 0.0;
        s.m2 = (ok && i >= 3 && i - 2 <= B) ? p[-2] : 0.0;
        s.m3 = (ok && i >= 4 && i - 3 <= B) ? p[-3] : 0.0;
        return s;
    };
    // one 4-row block: shuffle the four unresolved values, solve the 4x4 block redundantly, update
    // this lane's row; `nxt` is loaded for the following block while this one resolves
    auto step = [&](const SweepBlk &cur, SweepBlk &nxt, bool more) {
        const int i = (j - lane) & 31;
        const bool piv = i < 4;
        const int rn = piv ? r - 32 : r;
        nxt = load_blk(j - 4, rn, more);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double a0 = __shfl_sync(kFull, acc[v], j & 31), a1 = __shfl_sync(kFull, acc[v], (j - 1) & 31);
            const double a2 = __shfl_sync(kFull, acc[v], (j - 2) & 31), a3 = __shfl_sync(kFull, acc[v], (j - 3) & 31);
            const double fresh = (piv && rn >= 0) ? x[v * vs + rn] : 0.0;
            const double x0 = a0;
            const double x1 = fma(-cur.l10, x0, a1);
            const double x2 = fma(-cur.l21, x1, fma(-cur.l20, x0, a2));
            const double x3 = fma(-cur.l32, x2, fma(-cur.l31, x1, fma(-cur.l30, x0, a3)));
            double a = fma(-cur.m1, x1, fma(-cur.m0, x0, acc[v]));
            a = fma(-cur.m3, x3, fma(-cur.m2, x2, a));
            double xsel = x3;
            xsel = (i == 2) ? x2 : xsel;
            xsel = (i == 1) ? x1 : xsel;
            xsel = (i == 0) ? x0 : xsel;
            if (piv) x[v * vs + j - i] = xsel;
            acc[v] = piv ? fresh : a;
        }
        r = rn;
        j -= 4;
    };
    SweepBlk blkA = load_blk(j, r, nblk > 0), blkB;
    int b = 0;
#pragma unroll 1
    for (; b + 1 < nblk; b += 2) {  // two blocks per trip: the operand sets alternate, no register copies
        step(blkA, blkB, true);
        step(blkB, blkA, b + 2 < nblk);
    }
    if (b < nblk) step(blkA, blkB, false);
#pragma unroll 1
    for (; j >= lo; --j) {
        const int i = (j - lane) & 31;
        double lv = 0.0;
        if (i >= 1 && i <= B && r >= 0) lv = band[r * P + i];
        const bool piv = (i == 0);
        const int rn = piv ? r - 32 : r;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double xj = __shfl_sync(kFull, acc[v], j & 31);
            const double fresh = (piv && rn >= 0) ? x[v * vs + rn] : 0.0;
            if (piv) x[v * vs + j] = xj;
            acc[v] = piv ? fresh : fma(-lv, xj, acc[v]);
        }
        r = rn;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
        if (r >= 0) x[v * vs + r] = acc[v];  // rows (lo-32, lo): unfinished partial sums
    __syncwarp();
}

// Bottom front: the P middle unknowns are final (xmid points at the top front's middle rows,
// global order; bottom local row ncols + m <-> xmid[P-1-m]); fold them into the partial sums of
// rows [ncols-B, ncols) of NV vectors.
template <int B, int NV>
__device__ __forceinline__ void front_apply_known(const double *__restrict__ band, double *__restrict__ x,
                                                  const double *__restrict__ xmid, int vs, int ncols, int lane) {
    constexpr int P = B + 1;
    const int r = ncols - 1 - lane;
    if (lane < B && r >= 0) {
        const double *p = band + r * P;
        double a[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v] = x[v * vs + r];
#pragma unroll
        for (int o = 1; o <= B; ++o) {
            if (o > lane) {
                const double l = p[o];
                const int m = o - 1 - lane;  // local row ncols + m
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v] = fma(-l, xmid[v * vs + (P - 1 - m)], a[v]);
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) x[v * vs + r] = a[v];
    }
    __syncwarp();
}

}  // namespace vbfem
