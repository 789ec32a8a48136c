// vbfem_panel.cuh -- blocked banded LDL^T for wide bands (Cook 80x40: n = 6560, half bandwidth 85):
// one CTA per Monte-Carlo sample, two CTAs per SM, the factor streamed to HBM with bulk copies.
//
// The matrix is cut into 8x8 blocks.  A "panel" is one block column: the diagonal block and the
// NB blocks below it (NB = 11 for b = 85) plus ONE extra block row that carries up to eight
// right-hand sides through the elimination (row 0: the load vector, rows 1..6: the six strain
// functionals B^T of the two observed Gauss points).  Per panel:
//   diag   the 8x8 diagonal block is factored (LDL^T) and its unit lower factor inverted, in
//          registers, by one warp;
//   solve  every block below becomes V = X * L11^-T (two FP64 tensor-core MMAs m8n8k4 per block);
//   update the trailing window, (NB+1) NB/2 + NB blocks, takes its rank-8 update C -= L V^T, again
//          two DMMAs per block.  Fragments are read and written as ONE 16-byte access per lane at
//          (block base + 16 * lane): the 8x8 row-major block is exactly the C fragment layout,
//          and with the contraction index split {0,2,4,6} / {1,3,5,7} over the two MMAs it is
//          the A and B fragment layout too -- no shuffles, no bank conflicts;
//   fresh  the block row entering the window is assembled on the fly by an atomics-free GATHER
//          (host table: target entry <- up to four element-matrix entries) from a ring of element
//          matrices that the per-element Q4 kernels fill a batch ahead (src/mat_subroutine_tf.py:23-110,
//          src/fem_solver_tf.py:229-341 upstream).  K itself never exists in HBM.
// Only the lower triangle of the window is kept: diagonal d of the window is a ring of NB+2-d block
// slots (NB+1-d live blocks and one spare), the slot of block (I, J) is J mod (NB+2-d); 46 KB for NB = 11.
// The spare slot lets the block row entering the window be assembled WHILE the current panel is
// being applied.  Warp roles inside a panel step (two block barriers per panel):
//   solve   warps 0..5: the blocks below the (already factored) diagonal block
//   update  warps 0..5: the trailing blocks, software pipelined (the next block's fragments are in
//           flight while the current block's two MMAs run);
//           warp 6: updates the NEXT diagonal block first and factors it at once (look-ahead), and
//           sends the finished panel to HBM;  warp 7: element matrices + gather of the entering row.
//
// The band order ENDS at the observed node, so its displacement y falls out of the last diagonal
// block; the observed strains are eps_i = q_i^T K^-1 f = sum_c z_qi[c] z_f[c] / d_c, accumulated from
// the right-hand-side rows while they are eliminated: forward mode needs no back substitution and
// writes nothing to HBM.  With an adjoint (fused or Jacobian mode) the scaled panels (L^T blocks, the
// INVERSE of the diagonal block's unit factor, D^-1 z rows) go to a per-CTA HBM slab by
// cp.async.bulk, and ONE reverse pass (bulk loads through an mbarrier ring) back-substitutes u and
// up to four adjoint vectors together: the right-hand side of psi = K^-1 w is a combination of the
// stored rows, since w lies in the span of the strain functionals and the observed node.
// Replaces tf.linalg.solve (src/fem_solver_tf.py:137 upstream) and its gradient on wide bands.
#pragma once
#include "vbfem_front.cuh"

namespace vbfem {

constexpr int kPanelNT = 256, kPanelNW = kPanelNT / 32, kPanelNBMax = 15, kPanelStagesMax = 8;
constexpr int kPanelUpdW = 6;   // warps 0..5 solve and update; warp 6: diagonal look-ahead; warp 7: assembly
constexpr int kPanelEB = 8;     // element matrices per pass of one warp (lane = element x Gauss point)
constexpr int kPanelRecDepth = 4;  // row records in flight to the assembling warp (bulk copies into a shared-memory ring)

struct PanelModel {
    int n, off, npad, NQ, NB;  // order, leading pad rows, padded order, panels, block half bandwidth
    int R;                     // capacity of the element-matrix ring
    int nub;                   // blocks of one trailing update
    int obs_loc[2];            // row inside the last panel of the observed node's (x, y) dof, -1 if supported
    int o_win, o_rhs, o_lst, o_ke, smem_bytes;  // shared-memory offsets in bytes
    int stages;                // bulk-load ring of the reverse pass
    int o_lneg;                // NB+2 blocks: -L of the current panel, row major (block b: block row p+b; NB+1: rhs), then a dummy block
    int o_rec, rec_stride, rec_o_src, rec_o_dst;  // row-record ring in shared memory / record layout (bytes)
    // second generation (vbfem_panel2.cuh): shared-memory offsets of the diagonal ring, the entering-row staging area,
    // the V blocks, and of the region the reverse pass re-uses as its bulk-load ring
    int o_wdiag, o_fresh, o_vst, o_big;
    int kstart[kPanelNW + 1];  // update blocks [kstart[w], kstart[w+1]) belong to warp w < kPanelUpdW
    unsigned short ub[kPanelNBMax * (kPanelNBMax + 1) / 2 + kPanelNBMax];  // (I << 8) | J
    // Row record of block row q (rec_stride bytes, what the row needs to enter the window): int32 header
    // {new elements, gather entries, first new element (first-use order)}, the 8x8 right-hand-side block,
    // gather sources (ushort4: element-ring entries slot * 36 + tri, unused -> the zero entry) and gather
    // targets (u16: d * 64 + g * 8 + c)
    const unsigned char *rec;
    const int *eneed;            // [NQ] elements (first-use order) block row q needs
    const double *ecoord;        // [nele][4][2] nodal coordinates of the elements in first-use order
    double *kews;                // per-CTA scratch: the sample's element matrices [nele][36], first-use order
    long long kews_stride;
    const int *elm;              // [nele][8] padded band row of each element dof, -1 if supported
    double *lws;                 // per-CTA factor slab
    long long lws_stride;        // doubles
    double *xws;                 // per-CTA solution vectors [5][npad]
    long long xws_stride;
};

struct PanelSmem {
    double rd[2][8];    // 1/d of the diagonal block, by panel parity
    double minv[2][64]; // inverse of the diagonal block's unit factor, row major, by panel parity
    double W[64];       // [v][a]: weight of right-hand-side row a in the right-hand side of vector v
    double nodew[16];   // [v][2]: weight of the observed node's unit vectors
    double nodeL[16];   // [2][8]: D^-1 L11^-1 e_(observed dof) inside the last panel
    double G[8];        // q_a^T K^-1 f
    double lf_last[8];  // D^-1 z_f of the last panel
    double obs[32];
    double red[2 * 5 * kPanelNW];
    unsigned long long bar[kPanelStagesMax];
    unsigned long long rbar[kPanelRecDepth];
    uint4 utab[kPanelUpdW][32];  // per update warp: shared-memory byte offsets (A, B, C) of its blocks in this panel
    int colslot[2][kPanelNBMax + 2];
    int flag;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One 8x8x8 block product C += A * B^T on fragments (a = A[g][2t..2t+1], b = B[g][2t..2t+1], c = C[g][2t..2t+1]).
// DMMA: two tensor-core MMAs.  The alternative distributes the operands by shuffles and runs 16 DFMAs
// per lane -- the same arithmetic on the FP64 pipe, kept for the ncu comparison the design notes quote.
template <bool DMMA>
__device__ __forceinline__ void block_mma(double2 &c, const double2 a, const double2 b, int lane) {
    if (DMMA) {
        dmma884(c.x, c.y, a.x, b.x);
        dmma884(c.x, c.y, a.y, b.y);
    } else {
        const int g4 = lane & ~3, t = lane & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double ax = __shfl_sync(kFull, a.x, g4 + k), ay = __shfl_sync(kFull, a.y, g4 + k);
            const int r0 = 8 * t + k;  // lane holding B[2t][2k..2k+1], the next row is 4 lanes up
            const double b0x = __shfl_sync(kFull, b.x, r0), b0y = __shfl_sync(kFull, b.y, r0);
            const double b1x = __shfl_sync(kFull, b.x, r0 + 4), b1y = __shfl_sync(kFull, b.y, r0 + 4);
            c.x = fma(ax, b0x, c.x);
            c.x = fma(ay, b0y, c.x);
            c.y = fma(ax, b1x, c.y);
            c.y = fma(ay, b1y, c.y);
        }
    }
}

// LDL^T of an 8x8 diagonal block and the inverse of its unit factor in the MMA fragment layout, without redundant
// arithmetic (lane (g, t) holds D[g][2t], D[g][2t+1]; only the lower triangle of D is meaningful): right-looking LDL^T, column k broadcast by four shuffles per step (pivot, this
// lane's row entry, the entries of this lane's two columns), the inverse of the unit factor built alongside by forward
// substitution on the identity (row k of the inverse is final when step k starts).  80 FP64 instructions and ~100
// shuffles (the first version factored the block redundantly in every lane: 190 FP64 instructions, 18 broadcast loads,
// two warp barriers and a 36-entry working set in registers).  Returns Minv (C layout = the solve's B fragment), its transpose (the reverse
// pass's B fragment), the reciprocal pivots of this lane's two columns.
__device__ __forceinline__ void warp_diag_fragment(double2 D, double2 &Minv, double2 &MinvT, double2 &r2, int *flag,
                                                   int lane) {
    const int g = lane >> 2, t = lane & 3;
    double2 M = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);
    r2 = make_double2(0.0, 0.0);
    int bad = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int kt = k >> 1;
        const double src = (k & 1) ? D.y : D.x;  // column k lives in the lanes with t == k / 2
        const double dk = __shfl_sync(kFull, src, 4 * k + kt);
        const double cg = __shfl_sync(kFull, src, 4 * g + kt);
        const double cj0 = __shfl_sync(kFull, src, 8 * t + kt);
        const double cj1 = __shfl_sync(kFull, src, 8 * t + 4 + kt);
        bad |= (unsigned)(__double2hiint(dk) - 0x00200000) >= 0x7fd00000u;
        const double rk = fast_rcp3(dk);
        const double lg = (g > k) ? cg * rk : 0.0;  // L[g][k]
        if (2 * t > k) D.x = fma(-lg, cj0, D.x);
        if (2 * t + 1 > k) D.y = fma(-lg, cj1, D.y);
        if (t == kt) {
            if (k & 1) r2.y = rk;
            else r2.x = rk;
        }
        if (k < 7) {
            const double mkx = __shfl_sync(kFull, M.x, 4 * k + t), mky = __shfl_sync(kFull, M.y, 4 * k + t);
            M.x = fma(-lg, mkx, M.x);
            M.y = fma(-lg, mky, M.y);
        }
    }
    if (bad && lane == 0) *flag = 1;
    Minv = M;
    const int s0 = 8 * t + (g >> 1), s1 = s0 + 4;
    const double ax = __shfl_sync(kFull, M.x, s0), ay = __shfl_sync(kFull, M.y, s0);
    const double bx = __shfl_sync(kFull, M.x, s1), by = __shfl_sync(kFull, M.y, s1);
    MinvT = make_double2((g & 1) ? ay : ax, (g & 1) ? by : bx);
}

// The same factorisation with 2x2 PIVOT BLOCKS: columns (k, k+1) are eliminated together.  With a = D[k][k],
// b = D[k+1][k], c = D[k+1][k+1] and det = a c - b^2 (= a d_(k+1)), the trailing entries take
//     D[i][j] -= (p_i (c p_j - b q_j) + q_i (a q_j - b p_j)) / det,      p = D[.][k], q = D[.][k+1],
// where everything but 1 / det is computed while the reciprocal is in flight: the dependency chain of a PAIR of
// pivots is shuffle -> det (2 FP64) -> reciprocal -> one FMA, against two times shuffle -> reciprocal -> multiply ->
// FMA column by column.  A pivot pair sits in ONE set of lanes (t == k / 2, components x and y), so the shuffles
// stay the same in number.  Mathematically identical to two LDL^T steps (1 / d_k = 1 / a, 1 / d_(k+1) = a / det,
// L[i][k] = p_i / a, L[i][k+1] = (a q_i - b p_i) / det); the pivot check covers a and det (both must be positive).
__device__ __forceinline__ void warp_diag_fragment_pairs(double2 D, double2 &Minv, double2 &MinvT, double2 &r2, int *flag,
                                                         int lane) {
    const int g = lane >> 2, t = lane & 3;
    double2 M = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);
    r2 = make_double2(0.0, 0.0);
    int bad = 0;
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
        const int k = 2 * kp;
        const double a = __shfl_sync(kFull, D.x, 4 * k + kp);
        const double b = __shfl_sync(kFull, D.x, 4 * (k + 1) + kp);
        const double c = __shfl_sync(kFull, D.y, 4 * (k + 1) + kp);
        const double pg = __shfl_sync(kFull, D.x, 4 * g + kp), qg = __shfl_sync(kFull, D.y, 4 * g + kp);
        const double det = fma(a, c, -b * b);
        bad |= (unsigned)(__double2hiint(a) - 0x00200000) >= 0x7fd00000u;
        bad |= (unsigned)(__double2hiint(det) - 0x00200000) >= 0x7fd00000u;
        const double rdet = fast_rcp3(det);
        if (kp < 3) {
            const double p0 = __shfl_sync(kFull, D.x, 8 * t + kp), q0 = __shfl_sync(kFull, D.y, 8 * t + kp);
            const double p1 = __shfl_sync(kFull, D.x, 8 * t + 4 + kp), q1 = __shfl_sync(kFull, D.y, 8 * t + 4 + kp);
            const double ns0 = fma(c, p0, -b * q0), nt0 = fma(a, q0, -b * p0);
            const double ns1 = fma(c, p1, -b * q1), nt1 = fma(a, q1, -b * p1);
            const double w0 = fma(pg, ns0, qg * nt0), w1 = fma(pg, ns1, qg * nt1);
            if (t > kp) {
                D.x = fma(-w0, rdet, D.x);
                D.y = fma(-w1, rdet, D.y);
            }
        }
        const double ra = fast_rcp3(a);
        const double lgk = (g > k) ? pg * ra : 0.0;                            // L[g][k]
        const double lgk1 = (g > k + 1) ? fma(a, qg, -b * pg) * rdet : 0.0;    // L[g][k+1]
        if (t == kp) r2 = make_double2(ra, a * rdet);
        {
            const double mkx = __shfl_sync(kFull, M.x, 4 * k + t), mky = __shfl_sync(kFull, M.y, 4 * k + t);
            M.x = fma(-lgk, mkx, M.x);
            M.y = fma(-lgk, mky, M.y);
        }
        if (kp < 3) {
            const double mkx = __shfl_sync(kFull, M.x, 4 * (k + 1) + t), mky = __shfl_sync(kFull, M.y, 4 * (k + 1) + t);
            M.x = fma(-lgk1, mkx, M.x);
            M.y = fma(-lgk1, mky, M.y);
        }
    }
    if (bad && lane == 0) *flag = 1;
    Minv = M;
    const int s0 = 8 * t + (g >> 1), s1 = s0 + 4;
    const double ax = __shfl_sync(kFull, M.x, s0), ay = __shfl_sync(kFull, M.y, s0);
    const double bx = __shfl_sync(kFull, M.x, s1), by = __shfl_sync(kFull, M.y, s1);
    MinvT = make_double2((g & 1) ? ay : ax, (g & 1) ? by : bx);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

#ifdef VBFEM_TIMELINE
#define PTL_DECL long long ptl[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ptl_t = clock64()
#define PTL(i)                          \
    do {                                \
        const long long now_ = clock64(); \
        ptl[i] += now_ - ptl_t;         \
        ptl_t = now_;                   \
    } while (0)
#define PTL_FLUSH                                                                                  \
    do {                                                                                           \
        if (A.timeline && lane == 0 && (warp < 2 || warp >= 6))                                    \
            for (int i_ = 0; i_ < 16; ++i_)                                                        \
                A.timeline[(blockIdx.x * 4 + (warp < 2 ? warp : warp - 4)) * 16 + i_] = ptl[i_];   \
    } while (0)
#else
#define PTL_DECL ((void)0)
#define PTL(i) ((void)0)
#define PTL_FLUSH ((void)0)
#endif

// MODE 0: y, h   MODE 1: y, h, gx = J^T (gy, gh)   MODE 2: y, h, J = d(y, h)/dx
template <int MODE, bool DMMA>
__global__ void __launch_bounds__(kPanelNT, 2) fem_panel_kernel(const __grid_constant__ DevModel M,
                                                                const __grid_constant__ PanelModel Q,
                                                                const __grid_constant__ Args A) {
    extern __shared__ __align__(16) unsigned char smraw[];
    PanelSmem &S = *reinterpret_cast<PanelSmem *>(smraw);
    double *win = reinterpret_cast<double *>(smraw + Q.o_win);  // window blocks, then (contiguous) the rhs ring
    double *rhs = reinterpret_cast<double *>(smraw + Q.o_rhs);  // NB+2 right-hand-side blocks
    double *lst = reinterpret_cast<double *>(smraw + Q.o_lst);  // two staging panels (transposed, scaled)
    double *ke = reinterpret_cast<double *>(smraw + Q.o_ke);    // R element matrices (36 each), then 0.0, 1.0
    double *lneg = reinterpret_cast<double *>(smraw + Q.o_lneg);  // -L blocks of the current panel, then a dummy block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NB = Q.NB, NQ = Q.NQ, NB1 = NB + 1, NB2 = NB + 2, LPB = (NB + 2) * 64;  // LPB: doubles per stored panel
    constexpr int NV = (MODE == 2) ? 5 : 2;
    auto dbase = [&](int d) { return d * NB2 - (d * (d - 1)) / 2; };  // first slot of window diagonal d (ring of NB+2-d)
    auto wrap = [](int v, int m) { return v >= m ? v - m : v; };
    auto ldv = [&](unsigned off) { return reinterpret_cast<const double2 *>(smraw + off)[lane]; };
    double *lws = Q.lws + (size_t)blockIdx.x * Q.lws_stride;
    double *xws = Q.xws + (size_t)blockIdx.x * Q.xws_stride;
    // mbarriers (row records; reverse pass): initialised once, their phases run on across the CTA's samples
    if (tid == 0) {
        for (int i = 0; i < Q.stages; ++i) mbar_init(&S.bar[i], 1);
        for (int i = 0; i < kPanelRecDepth; ++i) mbar_init(&S.rbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned sweep_base = 0;  // bulk loads issued per stage ring so far (all samples of this CTA)
    unsigned rec_base = 0;    // row records loaded so far
    unsigned char *recs = smraw + Q.o_rec;

    for (long long s = blockIdx.x; s < A.N; s += gridDim.x) {
        // ---------------- sample parameters: theta -> (E, nu) -> (lambda, mu), recomputed where needed (three
        //                  places) instead of being carried through the panel loop in registers
        // src/data_generation_2sam_more_loss.py:181-186
        auto sample_material = [&](double &E, double &nu) {
            double x0, x1;
            if (A.mode & kElbo) {
                // main_custom_training.py:199-209: theta = e * sqrt(sig2) + mu, flattened [B*S]
                const long long j = A.j_begin + s;
                const int bb = (int)(j / A.S), ss = (int)(j % A.S);
                x0 = A.e[2 * ss] * sqrt(A.sig2[2 * bb]) + A.mu[2 * bb];
                x1 = A.e[2 * ss + 1] * sqrt(A.sig2[2 * bb + 1]) + A.mu[2 * bb + 1];
            } else {
                x0 = A.x[2 * s];
                x1 = A.x[2 * s + 1];
            }
            E = exp(M.theta_std[0] * x0 + M.theta_mean[0]);
            nu = 0.5 / (1.0 + exp(-M.theta_std[1] * x1 - M.theta_mean[1]));
            return lame_from_E_nu(E, nu);
        };
        PTL_DECL;

        // ---------------- reset: window and rhs ring to zero, ring constants, slot tables
        {
            double2 *w2 = reinterpret_cast<double2 *>(win);
            const int nz = (dbase(NB1) + NB2) * 32;  // window blocks + rhs ring, 32 double2 per block
            for (int i = tid; i < nz; i += kPanelNT) w2[i] = make_double2(0.0, 0.0);
            if (tid == 0) {
                ke[Q.R * 36] = 0.0;
                ke[Q.R * 36 + 1] = 1.0;
                S.flag = 0;
            }
            if (tid <= NB2) S.colslot[0][tid] = 0;
        }
        __syncthreads();

        double gacc = 0.0;  // warp 0: partial sum of G[g] over this lane's columns
        // (a) Per-element Q4 Gauss-point kernels of this sample, all warps, thread = element: shape functions,
        // material subroutine at the zero predictor, kt += dvol B^T Ct B over the 2x2 rule
        // (src/mat_subroutine_tf.py:23-110).  The 36 lower-triangle entries go to the CTA's scratch slab in
        // first-use order; the panel loop pulls them into the shared-memory ring a few rows ahead of their use.
        double *kews = Q.kews + (size_t)blockIdx.x * Q.kews_stride;
        {
        double E_, nu_;
        const Lame mat = sample_material(E_, nu_);
        for (int k = tid; k < M.nele; k += kPanelNT) {
            double xl[4], yl[4], kev[36];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double2 xy = reinterpret_cast<const double2 *>(Q.ecoord + (size_t)8 * k)[a];
                xl[a] = xy.x;
                yl[a] = xy.y;
            }
#pragma unroll
            for (int q = 0; q < 36; ++q) kev[q] = 0.0;
#pragma unroll 1
            for (int gp = 0; gp < 4; ++gp) {
                ShapeQ4 sh;
                shapef_q4(xl, yl, gp, M.thk, sh);
                double sig[4];
                Tangent C;
                mat_isotropic_plane_strain(mat, 0.0, 0.0, 0.0, sig, C);
                accumulate_kt(sh, C, kev);
            }
            double2 *dst = reinterpret_cast<double2 *>(kews + (size_t)36 * k);
#pragma unroll
            for (int q = 0; q < 18; ++q) dst[q] = make_double2(kev[2 * q], kev[2 * q + 1]);
        }
        }
        __threadfence_block();
        asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what was just written
        __syncthreads();
        // LDL^T of the diagonal block (p, p) and the inverse of its unit factor, by one warp: every lane
        // factors the 36 entries redundantly in registers (no exchange on the pivot chain), lane j < 8 then
        // forms column j of the inverse and stores it as row j of the transposed block stg[c][k] = Minv[k][c].
        auto diag_factor = [&](const double *D, double *stg, double *rdo, double *mro) {
            double a[36];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; j += 2) {
                    const double2 v = reinterpret_cast<const double2 *>(D + i * 8)[j >> 1];
                    a[tri(i, j)] = v.x;
                    if (j + 1 <= i) a[tri(i, j + 1)] = v.y;
                }
            double rdv[8];
            int bad = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = a[tri(k, k)];
                bad |= !(d > 0.0 && d < 1.0e300);
                rdv[k] = fast_rcp3(d);
#pragma unroll
                for (int j = k + 1; j < 8; ++j) {
                    const double ljk = a[tri(j, k)] * rdv[k];
#pragma unroll
                    for (int i = j; i < 8; ++i) a[tri(i, j)] = fma(-a[tri(i, k)], ljk, a[tri(i, j)]);
                    a[tri(j, k)] = ljk;  // rows i > j of column k stay unscaled until their own turn
                }
            }
            if (bad && lane == 0) S.flag = 1;
            const int j = lane & 7;
            double m[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) acc = fma(a[tri(i, k)], m[k], acc);  // m[k] = 0 for k < j
                m[i] = (i > j) ? -acc : m[i];
            }
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; i += 2)
                    reinterpret_cast<double2 *>(stg + j * 8)[i >> 1] = make_double2(m[i], m[i + 1]);
#pragma unroll
                for (int i = 0; i < 8; ++i) mro[i * 8 + j] = m[i];  // row-major copy: the solve's B fragments
            } else if (lane == 8) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) reinterpret_cast<double2 *>(rdo)[k >> 1] = make_double2(rdv[k], rdv[k + 1]);
            }
        };

        // ---------------- the first NB+1 block rows fill the window (all warps, records and element matrices
        //                  read from global; block (q, q-d) in slot q-d).  Meanwhile the records and element
        //                  matrices of the next rows are on their way into shared memory.
        const int nrec = (NQ > NB1) ? NQ - NB1 : 0;  // rows that enter during the panel loop
        auto fetch_row = [&](int q, int sl) {  // one thread: row record + the element matrices the row is first to need
            const int e0 = Q.eneed[q - 1], e1 = Q.eneed[q];
            mbar_expect_tx(&S.rbar[sl], Q.rec_stride + 288 * (e1 - e0));
            bulk_load(recs + sl * Q.rec_stride, Q.rec + (size_t)q * Q.rec_stride, Q.rec_stride, &S.rbar[sl]);
            for (int k = e0; k < e1; ++k) bulk_load(ke + (k % Q.R) * 36, kews + (size_t)36 * k, 288, &S.rbar[sl]);
        };
        {
            const int e1 = Q.eneed[NB < NQ ? NB : NQ - 1];
            for (int i = tid; i < e1 * 18; i += kPanelNT) {
                const int k = i / 18, j = i - 18 * k;
                reinterpret_cast<double2 *>(ke + (k % Q.R) * 36)[j] = reinterpret_cast<const double2 *>(kews + (size_t)36 * k)[j];
            }
        }
        __syncthreads();
        for (int q = 0; q <= NB && q < NQ; ++q) {
            const unsigned char *rc = Q.rec + (size_t)q * Q.rec_stride;
            const int4 hd = *reinterpret_cast<const int4 *>(rc);  // new elements, entries, first new element
            if (tid < 32) reinterpret_cast<double2 *>(rhs + q * 64)[tid] = reinterpret_cast<const double2 *>(rc + 16)[tid];
            const ushort4 *src = reinterpret_cast<const ushort4 *>(rc + Q.rec_o_src);
            const unsigned short *dstp = reinterpret_cast<const unsigned short *>(rc + Q.rec_o_dst);
            for (int i = tid; i < hd.y; i += kPanelNT) {
                const int dst = dstp[i];
                const ushort4 sr = src[i];
                const int d = dst >> 6;
                win[(dbase(d) + q - d) * 64 + (dst & 63)] = ((ke[sr.x] + ke[sr.y]) + ke[sr.z]) + ke[sr.w];
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 7 * 32)  // ring slots of elements the first rows no longer need may now be overwritten
            for (int j = 0; j < kPanelRecDepth && j < nrec; ++j) fetch_row(NB1 + j, (int)((rec_base + j) % kPanelRecDepth));
        if (warp == 6) diag_factor(win, lst, S.rd[0], S.minv[0]);  // block (0, 0): diagonal 0, slot 0
        __syncthreads();
        PTL(0);

        // ---------------- panels
        int rslot = 0;  // p mod (NB+2): slot of panel p in the rhs ring
        for (int p = 0; p < NQ; ++p) {
            const int par = p & 1;
            const int *cs = S.colslot[par];
            double *stg = lst + par * LPB;  // staging panel: [0] inverse unit factor^T, [1..NB] L^T blocks, [NB+1] rhs
            const double2 r2 = reinterpret_cast<const double2 *>(S.rd[par])[t];
            // ---- phase B: V = X L11^-T for the blocks below the diagonal block (in place) and the
            //      right-hand-side block; the scaled copy goes to the staging panel
            if (warp < kPanelUpdW) {
                const double2 mi = reinterpret_cast<const double2 *>(S.minv[par])[lane];  // Minv[g][2t..2t+1]
                const double2 idf = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);  // identity fragment
                for (int b = warp; b <= NB; b += kPanelUpdW) {  // b = 0: right-hand sides, else block row p+b
                    double2 *X = reinterpret_cast<double2 *>(b ? win + (dbase(b) + cs[b]) * 64 : rhs + rslot * 64);
                    const double2 xv = X[lane];
                    double2 v = make_double2(0.0, 0.0);
                    block_mma<DMMA>(v, xv, mi, lane);
                    X[lane] = v;
                    const double2 l = make_double2(v.x * r2.x, v.y * r2.y);
                    reinterpret_cast<double2 *>(lneg + (b ? b : NB + 1) * 64)[lane] = make_double2(-l.x, -l.y);
                    if (MODE > 0) {
                        // the stored panel holds L^T blocks: I * L^T on the tensor core leaves the transposed block in
                        // the (lane-contiguous) C fragment layout -- no bank-conflicted scatter
                        double2 lt = make_double2(0.0, 0.0);
                        block_mma<DMMA>(lt, idf, l, lane);
                        reinterpret_cast<double2 *>(stg + (b ? b : NB + 1) * 64)[lane] = lt;
                    }
                    if (b == 0) {  // warp 0: strain rows against the load row, G[g] += sum_c V[g][c] L[0][c]
                        const double lfx = __shfl_sync(kFull, l.x, t), lfy = __shfl_sync(kFull, l.y, t);
                        gacc = fma(v.x, lfx, fma(v.y, lfy, gacc));
                        if (p == NQ - 1 && g == 0) {
                            S.lf_last[2 * t] = l.x;
                            S.lf_last[2 * t + 1] = l.y;
                        }
                    }
                }
                if (MODE > 0) fence_async_smem();
            } else if (warp == 7 && lane <= NB2) {  // slot tables of panel p+1: (p+1) mod (NB+2-d)
                const int v = cs[lane] + 1;
                S.colslot[par ^ 1][lane] = (lane <= NB && v < NB2 - lane) ? v : 0;
            }
            PTL(1);
            __syncthreads();
            PTL(2);

            // ---- phase C
            if (warp < kPanelUpdW) {
                // trailing update C(I,J) -= L_I V_J^T.  Lane i works out the shared-memory offsets of this warp's
                // i-th block into a small table (one broadcast load per block in the loop).  One block per step, two
                // steps per trip on alternating register sets: the next block's fragments are fetched before the
                // current block's two (independent) MMAs are issued, and no register is ever copied.
                const int k0 = Q.kstart[warp], cnt = Q.kstart[warp + 1] - k0;
                if (lane < cnt) {
                    const int ub = Q.ub[k0 + lane], I = ub >> 8, J = ub & 255;
                    uint4 o;
                    o.x = Q.o_lneg + I * 512;
                    o.y = Q.o_win + (dbase(J) + cs[J]) * 512;
                    if (I <= NB) {
                        const int d = I - J;
                        o.z = Q.o_win + (dbase(d) + wrap(cs[d] + J, NB2 - d)) * 512;
                    } else {
                        o.z = Q.o_rhs + wrap(rslot + J, NB2) * 512;
                    }
                    o.w = 0;
                    S.utab[warp][lane] = o;
                }
                __syncwarp();
                const uint4 *tab = S.utab[warp];
                const unsigned lo = 16u * lane;
#define UPD_LOAD(X, P, k)  /* set X <- block k; its A fragment is taken from set P when the block row is the same */ \
    do {                                                                 \
        const uint4 t0_ = tab[k];                                        \
        X##c = t0_.z + lo;                                               \
        X##o = t0_.x;                                                    \
        X##a = P##a;                                                     \
        if (X##o != P##o) X##a = *reinterpret_cast<const double2 *>(smraw + t0_.x + lo); \
        X##b = *reinterpret_cast<const double2 *>(smraw + t0_.y + lo);   \
        X##v = *reinterpret_cast<const double2 *>(smraw + X##c);         \
    } while (0)
#define UPD_MMA_STORE(X)                                               \
    do {                                                               \
        if (DMMA) {                                                    \
            double2 e_ = make_double2(0.0, 0.0);                       \
            dmma884(X##v.x, X##v.y, X##a.x, X##b.x);                   \
            dmma884(e_.x, e_.y, X##a.y, X##b.y);                       \
            X##v.x += e_.x;                                            \
            X##v.y += e_.y;                                            \
        } else {                                                       \
            block_mma<false>(X##v, X##a, X##b, lane);                  \
        }                                                              \
        *reinterpret_cast<double2 *>(smraw + X##c) = X##v;             \
    } while (0)
                unsigned Xc, Yc, Xo, Yo = 0xffffffffu;
                double2 Xa, Xb, Xv, Ya = make_double2(0.0, 0.0), Yb, Yv;
                if (cnt > 0) UPD_LOAD(X, Y, 0);
                for (int i = 0; i < cnt; i += 2) {
                    if (i + 1 < cnt) UPD_LOAD(Y, X, i + 1);
                    UPD_MMA_STORE(X);
                    if (i + 1 >= cnt) break;
                    if (i + 2 < cnt) UPD_LOAD(X, Y, i + 2);
                    UPD_MMA_STORE(Y);
                }
#undef UPD_LOAD
#undef UPD_MMA_STORE
            } else if (warp == 6) {
                // the finished panel leaves for HBM; then block (p+1, p+1) gets its update ahead of the others and
                // is factored at once, so that the next panel's solve can start right after the barrier
                if (MODE > 0 && lane == 0) {
                    bulk_store(lws + (size_t)p * LPB, stg, LPB * 8);
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // panel p-1 has left the other buffer
                }
                if (p + 1 < NQ) {
                    __syncwarp();
                    const double2 v1 = reinterpret_cast<const double2 *>(win + (dbase(1) + cs[1]) * 64)[lane];
                    double *Dn = win + (dbase(0) + wrap(cs[0] + 1, NB2)) * 64;
                    double2 c = reinterpret_cast<double2 *>(Dn)[lane];
                    block_mma<DMMA>(c, reinterpret_cast<const double2 *>(lneg + 64)[lane], v1, lane);
                    reinterpret_cast<double2 *>(Dn)[lane] = c;
                    __syncwarp();
                    diag_factor(Dn, lst + (par ^ 1) * LPB, S.rd[par ^ 1], S.minv[par ^ 1]);
                }
            } else {
                // block row q = p+NB+1 enters the window through the spare slots (block (q, q-d): slot of
                // relative column NB+1-d): clear, right-hand-side block, element matrices, gather -- all from
                // the row's record, which a bulk copy brought into shared memory several panels ago
                const int q = p + NB1;
                const double2 z2 = make_double2(0.0, 0.0);
                for (int d = 0; d <= NB; ++d)
                    reinterpret_cast<double2 *>(win + (dbase(d) + wrap(cs[d] + NB1 - d, NB2 - d)) * 64)[lane] = z2;
                double2 *rdst = reinterpret_cast<double2 *>(rhs + wrap(rslot + NB1, NB2) * 64);
                if (q < NQ) {
                    const unsigned use = rec_base + (unsigned)p;
                    const int sl = (int)(use % kPanelRecDepth);
                    const unsigned char *rc = recs + sl * Q.rec_stride;
                    mbar_wait(&S.rbar[sl], (use / kPanelRecDepth) & 1u);
                    const int4 hd = *reinterpret_cast<const int4 *>(rc);
                    rdst[lane] = reinterpret_cast<const double2 *>(rc + 16)[lane];
                    const ushort4 *src = reinterpret_cast<const ushort4 *>(rc + Q.rec_o_src);
                    const unsigned short *dstp = reinterpret_cast<const unsigned short *>(rc + Q.rec_o_dst);
                    for (int i = lane; i < hd.y; i += 32) {
                        const int dst = dstp[i];
                        const ushort4 sr = src[i];
                        const int d = dst >> 6;
                        win[(dbase(d) + wrap(cs[d] + NB1 - d, NB2 - d)) * 64 + (dst & 63)] =
                            ((ke[sr.x] + ke[sr.y]) + ke[sr.z]) + ke[sr.w];
                    }
                    __syncwarp();
                    if (lane == 0 && p + kPanelRecDepth < nrec) fetch_row(q + kPanelRecDepth, sl);  // this slot's next tenant
                } else {
                    rdst[lane] = z2;
                }
            }
            PTL(3);
            __syncthreads();
            PTL(4);
            rslot = (rslot + 1 == NB2) ? 0 : rslot + 1;
        }

        // ---------------- observations: y from the last diagonal block, strains from the accumulated
        //                  products, h = von Mises at the two observed Gauss points (src/fem_postprocess.py:172-185)
        const double *stgl = lst + ((NQ - 1) & 1) * LPB;  // last panel: [c][k] = Minv[k][c]
        const double *rdl = S.rd[(NQ - 1) & 1];
        if (MODE > 0 && tid == 6 * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (warp == 0) {
            gacc += __shfl_xor_sync(kFull, gacc, 1);
            gacc += __shfl_xor_sync(kFull, gacc, 2);
            if (t == 0) S.G[g] = gacc;
            // D^-1 L11^-1 e_j for the observed node's dofs j (their unit vectors start in the last panel)
            if (lane < 16) {
                const int k = lane >> 3, c = lane & 7, j = Q.obs_loc[k];
                S.nodeL[lane] = (j >= 0) ? stgl[j * 8 + c] * rdl[c] : 0.0;
            }
        }
        __syncthreads();
        if (tid < 2) {
            double E_, nu_;
            const Lame mat = sample_material(E_, nu_);
            double exx = S.G[1 + 3 * tid], eyy = S.G[2 + 3 * tid], gxy = S.G[3 + 3 * tid];
            double sig[4];
            Tangent C;
            mat_isotropic_plane_strain(mat, exx, eyy, gxy, sig, C);
            double ds[4];
            const double hv = von_mises_ref(sig, ds);
            const double l2m = mat.lam + 2.0 * mat.mu;
            double *o = S.obs + 8 * tid;
            o[0] = hv;
            o[1] = ds[0] * l2m + ds[1] * mat.lam + ds[2] * mat.lam;  // dh/d(exx)
            o[2] = ds[0] * mat.lam + ds[1] * l2m + ds[2] * mat.lam;  // dh/d(eyy)
            o[3] = ds[3] * mat.mu;                                   // dh/d(gxy)
            o[4] = (ds[0] + ds[1] + ds[2]) * (exx + eyy);            // dh/d(lambda) at fixed u
            o[5] = 2.0 * ds[0] * exx + 2.0 * ds[1] * eyy + ds[3] * gxy;
            if (A.h) A.h[2 * s + tid] = hv;
            // y_k = (L11^-T D^-1 z_f)[j] = sum_c Minv[c][j] lf[c]
            const int j = Q.obs_loc[tid];
            double yv = 0.0;
            if (j >= 0)
                for (int c = 0; c < 8; ++c) yv = fma(stgl[j * 8 + c], S.lf_last[c], yv);
            S.obs[16 + tid] = yv;
            if (A.y) A.y[2 * s + tid] = yv;
            if (A.f_out) A.f_out[2 * s + tid] = yv;
            if (!(fabs(yv) < 1.0e300) || !(hv < 1.0e300)) S.flag = 1;
        }
        if (MODE > 0) {
            __syncthreads();
            // ---------------- right-hand sides of the reverse pass: v = 0 is u (row 0 = D^-1 z_f); the adjoint
            //                  vectors combine the strain rows and the observed node's unit vectors
            if (tid < 64) S.W[tid] = 0.0;
            if (tid < 16) S.nodew[tid] = 0.0;
            __syncthreads();
            if (tid == 0) {
                S.W[0] = 1.0;
                if (MODE == 1) {
                    double gy0, gy1, gh0 = 0.0, gh1 = 0.0;
                    if (A.mode & kElbo) {
                        // d(loss)/d f_j through term2 with the [B, B*S] broadcast (main_custom_training.py:205-214)
                        gy0 = A.gcoef * ((double)A.B * S.obs[16] - A.ysum[0]);
                        gy1 = A.gcoef * ((double)A.B * S.obs[17] - A.ysum[1]);
                    } else {
                        gy0 = A.gy[2 * s];
                        gy1 = A.gy[2 * s + 1];
                        gh0 = A.gh[2 * s];
                        gh1 = A.gh[2 * s + 1];
                    }
                    S.obs[20] = gh0;
                    S.obs[21] = gh1;
                    for (int i = 0; i < 3; ++i) {
                        S.W[8 + 1 + i] = gh0 * S.obs[1 + i];
                        S.W[8 + 4 + i] = gh1 * S.obs[8 + 1 + i];
                    }
                    S.nodew[2] = gy0;
                    S.nodew[3] = gy1;
                } else {
                    // vectors 1, 2: adjoints of y0, y1; 3, 4: adjoints of h0, h1
                    S.nodew[2 * 1] = 1.0;
                    S.nodew[2 * 2 + 1] = 1.0;
                    for (int i = 0; i < 3; ++i) {
                        S.W[3 * 8 + 1 + i] = S.obs[1 + i];
                        S.W[4 * 8 + 4 + i] = S.obs[8 + 1 + i];
                    }
                }
            }
            // ---------------- reverse pass: x_p = Minv_p^T (W Lrhs_p - sum_d x_(p+d) L_(p+d,p)), panels descending.
            //   Warp 0 finishes panel p (its products with x_(p+2..) were formed one step earlier), warps 1..7
            //   form the products of panel p-1 with the blocks that are already final: one barrier per panel.
            double *stage0 = win;                 // bulk-load ring (window + rhs ring are free now)
            double *xr = ke;                      // NB+1 solution blocks [v][k]
            double *part = ke + NB1 * 64;         // [2][8] partial products, by panel parity
            const int NS = Q.stages;
            for (int i = tid; i < (NB1 + 2 * kPanelNW) * 64; i += kPanelNT) xr[i] = 0.0;
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                for (int i = 0; i < NS && i < NQ; ++i) {
                    const int st = (sweep_base + i) % NS;
                    mbar_expect_tx(&S.bar[st], LPB * 8);
                    bulk_load(stage0 + st * LPB, lws + (size_t)(NQ - 1 - i) * LPB, LPB * 8, &S.bar[st]);
                }
            }
            PTL(7);
            int xs = (NQ - 1) % NB1;  // slot of panel p in the solution ring
            for (int i = 0; i < NQ; ++i) {
                const int p = NQ - 1 - i;
                const unsigned use = sweep_base + i;
                if (warp == 0) {
                    const int st = (int)(use % NS);
                    const double *pan = stage0 + st * LPB;
                    mbar_wait(&S.bar[st], (use / NS) & 1u);
                    double2 c = make_double2(0.0, 0.0), c2 = c;
                    block_mma<DMMA>(c, reinterpret_cast<const double2 *>(S.W)[lane],
                                    reinterpret_cast<const double2 *>(pan + (NB + 1) * 64)[lane], lane);
                    {
                        double2 a = reinterpret_cast<const double2 *>(xr + wrap(xs + 1, NB1) * 64)[lane];
                        a.x = -a.x;
                        a.y = -a.y;
                        block_mma<DMMA>(c2, a, reinterpret_cast<const double2 *>(pan + 64)[lane], lane);
                    }
                    double2 d = make_double2(c.x + c2.x, c.y + c2.y);
#pragma unroll
                    for (int w = 1; w < kPanelNW; ++w) {
                        const double2 q = reinterpret_cast<const double2 *>(part + ((p & 1) * kPanelNW + w) * 64)[lane];
                        d.x += q.x;
                        d.y += q.y;
                    }
                    if (p == NQ - 1) {
                        const double w0 = S.nodew[2 * g], w1 = S.nodew[2 * g + 1];
                        d.x += w0 * S.nodeL[2 * t] + w1 * S.nodeL[8 + 2 * t];
                        d.y += w0 * S.nodeL[2 * t + 1] + w1 * S.nodeL[8 + 2 * t + 1];
                    }
                    const double2 mi = reinterpret_cast<const double2 *>(pan)[lane];  // [c][k] = Minv[k][c]
                    double2 x = make_double2(0.0, 0.0);
                    block_mma<DMMA>(x, d, mi, lane);
                    reinterpret_cast<double2 *>(xr + xs * 64)[lane] = x;
                    if (g < NV) *reinterpret_cast<double2 *>(xws + (size_t)g * Q.npad + 8 * p + 2 * t) = x;
                } else if (p > 0) {
                    const int st = (int)((use + 1) % NS);
                    const double *pan = stage0 + st * LPB;
                    mbar_wait(&S.bar[st], ((use + 1) / NS) & 1u);
                    double2 c = make_double2(0.0, 0.0);
                    for (int b = 1 + warp; b <= NB; b += kPanelNW - 1) {  // block rows (p-1)+b, b >= 2
                        double2 a = reinterpret_cast<const double2 *>(xr + wrap(xs + b - 1, NB1) * 64)[lane];
                        a.x = -a.x;
                        a.y = -a.y;
                        block_mma<DMMA>(c, a, reinterpret_cast<const double2 *>(pan + b * 64)[lane], lane);
                    }
                    reinterpret_cast<double2 *>(part + (((p - 1) & 1) * kPanelNW + warp) * 64)[lane] = c;
                }
                __syncthreads();
                if (tid == 0 && i + NS < NQ) {
                    const int st = (int)(use % NS);
                    mbar_expect_tx(&S.bar[st], LPB * 8);
                    bulk_load(stage0 + st * LPB, lws + (size_t)(p - NS) * LPB, LPB * 8, &S.bar[st]);
                }
                xs = (xs == 0) ? NB : xs - 1;
            }
            sweep_base += (unsigned)NQ;
            PTL(8);

            // ---------------- element-wise contraction -psi^T (dK/dp) u + explicit dh/dp, chained to x
            constexpr int NADJ = NV - 1;
            double sl[NADJ], sm[NADJ];
#pragma unroll
            for (int v = 0; v < NADJ; ++v) sl[v] = sm[v] = 0.0;
            for (int e = tid; e < M.nele; e += kPanelNT) {
                double xl[4], yl[4], ue[8];
                int lm[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int nd = M.ien[4 * e + a];
                    const double2 xy = *reinterpret_cast<const double2 *>(M.coord + 2 * nd);
                    xl[a] = xy.x;
                    yl[a] = xy.y;
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    lm[a] = Q.elm[8 * e + a];
                    ue[a] = (lm[a] >= 0) ? xws[lm[a]] : 0.0;
                }
#pragma unroll 1
                for (int gp = 0; gp < 4; ++gp) {
                    ShapeQ4 sh;
                    shapef_q4(xl, yl, gp, M.thk, sh);
                    double uxx, uyy, uxy;
                    strain_q4(sh, ue, uxx, uyy, uxy);
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double *pv = xws + (size_t)(v + 1) * Q.npad;
                        double pe[8], pxx, pyy, pxy, cl, cm;
#pragma unroll
                        for (int a = 0; a < 8; ++a) pe[a] = (lm[a] >= 0) ? pv[lm[a]] : 0.0;
                        strain_q4(sh, pe, pxx, pyy, pxy);
                        mat_tangent_param_contract(pxx, pyy, pxy, uxx, uyy, uxy, cl, cm);
                        sl[v] = fma(sh.dvol, cl, sl[v]);
                        sm[v] = fma(sh.dvol, cm, sm[v]);
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < NADJ; ++v) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sl[v] += __shfl_down_sync(kFull, sl[v], o);
                    sm[v] += __shfl_down_sync(kFull, sm[v], o);
                }
                if (lane == 0) {
                    S.red[2 * (v * kPanelNW + warp)] = sl[v];
                    S.red[2 * (v * kPanelNW + warp) + 1] = sm[v];
                }
            }
            __syncthreads();
            if (tid == 0) {
                // d lambda, d mu / d(E, nu), then dE/dx0 = std0 * E ; dnu/dx1 = std1 * nu (1 - 2 nu)
                double E, nu;
                const Lame mat = sample_material(E, nu);
                const double tt = (1.0 + nu) * (1.0 - 2.0 * nu);
                const double dl_dE = mat.lam / E, dm_dE = mat.mu / E;
                const double dl_dnu = E * (1.0 + 2.0 * nu * nu) / (tt * tt);
                const double dm_dnu = -0.5 * E / ((1.0 + nu) * (1.0 + nu));
                const double dE_dx0 = M.theta_std[0] * E, dnu_dx1 = M.theta_std[1] * nu * (1.0 - 2.0 * nu);
                double tl[NADJ], tm[NADJ];
#pragma unroll
                for (int v = 0; v < NADJ; ++v) {
                    tl[v] = tm[v] = 0.0;
                    for (int w = 0; w < kPanelNW; ++w) {
                        tl[v] += S.red[2 * (v * kPanelNW + w)];
                        tm[v] += S.red[2 * (v * kPanelNW + w) + 1];
                    }
                }
                if (MODE == 1) {
                    const double gh0 = S.obs[20], gh1 = S.obs[21];
                    const double gl = -tl[0] + gh0 * S.obs[4] + gh1 * S.obs[8 + 4];
                    const double gm = -tm[0] + gh0 * S.obs[5] + gh1 * S.obs[8 + 5];
                    A.gx[2 * s] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                    A.gx[2 * s + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                } else {
                    // adjoint vectors v = 0, 1: y0, y1; v = 2, 3: h0, h1 -- the storage order of J
                    double *J = A.ws + (size_t)s * A.ws_stride;
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double gl = -tl[v] + (v >= 2 ? S.obs[8 * (v - 2) + 4] : 0.0);
                        const double gm = -tm[v] + (v >= 2 ? S.obs[8 * (v - 2) + 5] : 0.0);
                        J[2 * v] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                        J[2 * v + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && A.status) A.status[s] = S.flag;
        rec_base += (unsigned)nrec;
        __syncthreads();
        PTL(9);
        PTL_FLUSH;
    }
}

}  // namespace vbfem
