// vbfem_warp.cuh -- narrow bands (Cook 20x10: n = 440, half bandwidth 25): ONE WARP per Monte-Carlo sample,
// a dozen samples in flight per SM, the whole elimination window in REGISTERS.
//
// Same mathematics as the blocked panel kernel (vbfem_panel.cuh): 8x8 blocks, per panel the diagonal block is
// factored (LDL^T) and its unit factor inverted, the blocks below become V = X L11^-T (two FP64 tensor-core
// MMAs m8n8k4 per block) and the trailing window takes C -= L V^T (two more per block).  What is different:
//   * the window -- (NB+1)(NB+2)/2 = 10 blocks for NB = 3, plus NB+1 right-hand-side blocks -- never leaves
//     the register file: an 8x8 block in the MMA's C-fragment layout (lane (g, t) holds [g][2t], [g][2t+1]) IS
//     the A and the B fragment of the next product when the contraction index is split {0,2,4,6} / {1,3,5,7},
//     so solve -> scale -> update chain from registers to registers;
//   * the window slides for free: the update of block (I, J) is written to the registers of block
//     (I-1, J-1) (the MMA's D operand), nothing is copied;
//   * there is no block barrier and no cross-warp traffic in the sample loop.  Warps of a CTA share nothing
//     but the read-only tables; each pulls its next sample from a shared-memory counter;
//   * shared memory per warp is the ring of element matrices (just-in-time batches of 32, lane = element:
//     the per-element Q4 Gauss-point kernels of src/mat_subroutine_tf.py:23-110 upstream), the staging area
//     of the block row that enters the window (atomics-free gather, same host table as the panel kernel) and
//     a 1.5 KB exchange area for the diagonal block: ~17 KB, so 12-13 samples are resident per SM where the
//     on-chip two-front kernel (97 KB of factor per sample) holds two.  The latency of one sample's pivot
//     chain is hidden by the other samples; DMMA work (tensor pipe) and the diagonal-block factorisation
//     (FP64 pipe) of different warps overlap.
// With an adjoint (fused or Jacobian mode) the scaled panels leave for a per-warp slab in global memory as
// plain 16-byte stores from the fragments and come back, fragment by fragment, in ONE reverse pass that
// back-substitutes u and the adjoint vectors together -- again registers only.
// Replaces tf.linalg.solve (src/fem_solver_tf.py:137 upstream) and its gradient.
#pragma once
#include "vbfem_panel.cuh"

namespace vbfem {

constexpr int kWarpNB = 3;      // block half bandwidth of the register window (8x8 blocks)
constexpr int kWarpBatch = 32;  // element matrices per just-in-time batch (lane = element); WarpModel::batch may be smaller
constexpr int kWarpFixed = 640 + (kWarpNB + 1) * 512;  // bytes per warp ahead of the element ring: Minv^T and 1/d of the last panel, flag, staging area

struct WarpModel {
    int n, off, npad, NQ, R, nele;
    int batch;                   // element matrices per just-in-time batch (<= 32: lanes >= batch idle)
    int obs_loc[2];              // row inside the last panel of the observed node's (x, y) dof, -1 if supported
    int warp_smem;               // bytes of shared memory per warp
    int nent, tab_bytes;         // gather entries; bytes of the CTA-shared tables ahead of the per-warp areas
    // Gather table, one 64-bit word per target entry of the lower band: bits [0, 44) four 11-bit element-ring
    // entries slot * 36 + tri (unused: the zero entry), bits [44, 52) the target d * 64 + g * 8 + c (d: block
    // diagonal).  Row table: {first gather entry of block row q, elements (first-use order) the row needs
    // | 1 << 30 if the row has a non-zero initial right-hand-side block}.  Both are copied to shared memory once
    // per CTA.
    const unsigned long long *gpack;
    const int2 *rowtab;          // [NQ+1]
    const double *rhs0;          // [NQ][64] initial right-hand-side blocks [a][c]
    const double *ecoord;        // [nele][4][2] nodal coordinates, first-use order
    const int *elm;              // [nele][8] padded band row of each element dof, -1 if supported; first-use order
    int x_in_smem;               // fused adjoint: u and psi live in the (then idle) element ring instead of xws
    double *lws;                 // per-warp factor slab [NQ][NB+2][64]
    long long lws_stride;
    double *xws;                 // per-warp solution vectors [5][npad]
    long long xws_stride;
    // second generation (vbfem_warp2.cuh): K = lambda K_lam + mu K_mu from a sample-independent band table
    const double2 *ktab;         // (K_lam, K_mu)[npad][ldt], entry (r, c) at [r][r - c]; copied to shared memory per CTA
    int ldt, hb;                 // row stride (even, >= hb + 1), half bandwidth in padded rows
    int win_rows;                // ring of u / psi rows per warp (reverse pass): 40, or hb + 8 when shared memory is tight
    unsigned cmagic;             // id / (hb + 1) == (id * cmagic) >> 16 for id < 8 (hb + 1)
    unsigned long long rhsmask[2];  // bit q: block row q has a non-zero initial right-hand-side block
    int rhs_first;               // first such block row (NQ if none): the right-hand-side rows are zero above it
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One just-in-time batch of element matrices (lane = element, first-use order) into the ring.  Kept out of line:
// the 36 accumulators and the shape-function state would otherwise sit on top of the register window.
__device__ __noinline__ void warp_element_batch(const double *__restrict__ ecoord, double *__restrict__ ke, int k,
                                                int nele, int R, double thk, double lam, double mu) {
    if (k >= nele) return;
    Lame mat;
    mat.lam = lam;
    mat.mu = mu;
    mat.szz = lam;
    double xl[4], yl[4], kev[36];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double2 xy = reinterpret_cast<const double2 *>(ecoord + (size_t)8 * k)[a];
        xl[a] = xy.x;
        yl[a] = xy.y;
    }
#pragma unroll
    for (int q = 0; q < 36; ++q) kev[q] = 0.0;
#pragma unroll 1
    for (int gp = 0; gp < 4; ++gp) {
        ShapeQ4 sh;
        shapef_q4(xl, yl, gp, thk, sh);
        double sig[4];
        Tangent C;
        mat_isotropic_plane_strain(mat, 0.0, 0.0, 0.0, sig, C);
        accumulate_kt(sh, C, kev);
    }
    double2 *dst = reinterpret_cast<double2 *>(ke + (k % R) * 36);
#pragma unroll
    for (int q = 0; q < 18; ++q) dst[q] = make_double2(kev[2 * q], kev[2 * q + 1]);
}

#ifdef VBFEM_TIMELINE
#define WTL_DECL long long wtl[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, wtl_t = clock64()
#define WTL(i)                            \
    do {                                  \
        const long long now_ = clock64(); \
        wtl[i] += now_ - wtl_t;           \
        wtl_t = now_;                     \
    } while (0)
#define WTL_FLUSH                                                                   \
    do {                                                                            \
        if (A.timeline && lane == 0 && warp < 4)                                    \
            for (int i_ = 0; i_ < 16; ++i_) A.timeline[(blockIdx.x * 4 + warp) * 16 + i_] = wtl[i_]; \
    } while (0)
#else
#define WTL_DECL ((void)0)
#define WTL(i) ((void)0)
#define WTL_FLUSH ((void)0)
#endif

// MODE 0: y, h   MODE 1: y, h, gx = J^T (gy, gh)   MODE 2: y, h, J = d(y, h)/dx
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) fem_warp_kernel(const __grid_constant__ DevModel M,
                                                              const __grid_constant__ WarpModel Q,
                                                              const __grid_constant__ Args A) {
    constexpr int NB = kWarpNB, NB1 = NB + 1, LPB = (NB + 2) * 64;
    constexpr int NV = (MODE == 2) ? 5 : 2;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ int next_i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const unsigned long long *gtab = reinterpret_cast<const unsigned long long *>(smraw);
    const int2 *rowtab = reinterpret_cast<const int2 *>(smraw + (size_t)Q.nent * 8);
    unsigned char *wsm = smraw + Q.tab_bytes + (size_t)warp * Q.warp_smem;
    double *stg = reinterpret_cast<double *>(wsm), *rd = stg + 64;  // the last panel's Minv^T and 1 / d (observations)
    int *flagp = reinterpret_cast<int *>(wsm + 576);
    double *stage = reinterpret_cast<double *>(wsm + 640);  // NB+1 blocks of the entering row, by block diagonal
    double *ke = reinterpret_cast<double *>(wsm + kWarpFixed);  // R element matrices (36 each), then 0.0, 1.0
    // after the forward pass the staging area holds the small vectors of the observation / reverse pass
    double *sW = stage, *nodew = stage + 64, *nodeL = stage + 80, *sG = stage + 96, *lf_last = stage + 104,
           *obs = stage + 112;
    const int NQ = Q.NQ;
    const int wid = blockIdx.x * NW + warp;
    double *lws = Q.lws + (size_t)wid * Q.lws_stride;
    double *xws = Q.xws + (size_t)wid * Q.xws_stride;
    const double2 z2 = make_double2(0.0, 0.0);
    if (threadIdx.x == 0) next_i = 0;
    for (int i = threadIdx.x; i < Q.nent; i += NW * 32) reinterpret_cast<unsigned long long *>(smraw)[i] = Q.gpack[i];
    for (int i = threadIdx.x; i <= Q.NQ; i += NW * 32) reinterpret_cast<int2 *>(smraw + (size_t)Q.nent * 8)[i] = Q.rowtab[i];
    if (lane == 0) {
        ke[Q.R * 36] = 0.0;
        ke[Q.R * 36 + 1] = 1.0;
    }
    __syncthreads();

    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&next_i, 1);
        it = __shfl_sync(kFull, it, 0);
        const long long s = blockIdx.x + (long long)gridDim.x * it;
        if (s >= A.N) break;

        // ---------------- sample parameters: theta -> (E, nu) -> (lambda, mu)  (src/data_generation_2sam_more_loss.py:181-186)
        double E_, nu_;
        {
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
            E_ = exp(M.theta_std[0] * x0 + M.theta_mean[0]);
            nu_ = 0.5 / (1.0 + exp(-M.theta_std[1] * x1 - M.theta_mean[1]));
        }
        const Lame mat = lame_from_E_nu(E_, nu_);
        if (lane == 0) *flagp = 0;
        WTL_DECL;

        int computed = 0;  // element matrices (first-use order) in the ring so far
        // Block row q enters the window in two steps around a warp barrier: (a) the element matrices the row is the
        // first to need (just-in-time batch) and a cleared staging area, (b) the atomics-free gather of the row's
        // lower-band entries from the ring (at most 96 per row: three predicated trips, no loop, so that the
        // scheduler can interleave them with the diagonal block's factorisation).
        auto row_prepare = [&](int q) {
            const int need = rowtab[q].y & 0x3fffffff;
            while (computed < need) {
                warp_element_batch(Q.ecoord, ke, lane < Q.batch ? computed + lane : Q.nele, Q.nele, Q.R, M.thk, mat.lam,
                                   mat.mu);
                computed += Q.batch;
            }
            double2 *st2 = reinterpret_cast<double2 *>(stage);
#pragma unroll
            for (int i = 0; i < NB1; ++i) st2[i * 32 + lane] = z2;
        };
        auto row_gather = [&](int q) {
            const int e0 = rowtab[q].x, e1 = rowtab[q + 1].x;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int i = e0 + lane + 32 * r;
                if (i < e1) {
                    const unsigned long long w = gtab[i];
                    const unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
                    const double v = ((ke[lo & 2047u] + ke[(lo >> 11) & 2047u]) + ke[(unsigned)(w >> 22) & 2047u]) +
                                     ke[(hi >> 1) & 2047u];
                    stage[(hi >> 12) & 255u] = v;
                }
            }
        };

        // ---------------- the window: block (p+I, p+J) in W[I][J] (J <= I), right-hand sides of block column p+J in Rh[J]
        double2 W[NB1][NB1], Rh[NB1];
#pragma unroll
        for (int I = 0; I < NB1; ++I) {
            Rh[I] = z2;
#pragma unroll
            for (int J = 0; J < NB1; ++J) W[I][J] = z2;
        }
#pragma unroll
        for (int q = 0; q < NB1; ++q) {
            row_prepare(q);
            __syncwarp();
            row_gather(q);
            __syncwarp();
#pragma unroll
            for (int d = 0; d <= q; ++d) W[q][q - d] = reinterpret_cast<const double2 *>(stage)[d * 32 + lane];
            if (rowtab[q].y >> 30) Rh[q] = __ldg(reinterpret_cast<const double2 *>(Q.rhs0 + (size_t)q * 64) + lane);
            __syncwarp();
        }

        WTL(0);
        // ---------------- panels
        double gacc = 0.0;  // partial sum over this lane's columns of G[g] = q_g^T K^-1 f
        const double2 idf = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);  // identity fragment
        double2 mi, mit, r2;  // of the current panel: Minv[g][2t..2t+1], its transpose, 1 / d of columns 2t, 2t+1
        warp_diag_fragment(W[0][0], mi, mit, r2, flagp, lane);  // diagonal block of panel 0
#pragma unroll 1
        for (int p = 0; p < NQ; ++p) {
            WTL(1);
            // ---- solve: V = X L11^-T for the right-hand sides (b = 0) and the blocks below; Ln = -V D^-1
            double2 V[NB1], Ln[NB1];
#pragma unroll
            for (int b = 0; b < NB1; ++b) {
                const double2 xv = b ? W[b][0] : Rh[0];
                double2 v = z2;
                block_mma<true>(v, xv, mi, lane);
                V[b] = v;
                Ln[b] = make_double2(-v.x * r2.x, -v.y * r2.y);
            }
            {  // strain rows against the load row: G[g] += sum_c V[g][c] L[0][c]
                const double lfx = __shfl_sync(kFull, Ln[0].x, t), lfy = __shfl_sync(kFull, Ln[0].y, t);
                gacc = fma(V[0].x, -lfx, fma(V[0].y, -lfy, gacc));
                if (p == NQ - 1 && g == 0) {
                    lf_last[2 * t] = -Ln[0].x;
                    lf_last[2 * t + 1] = -Ln[0].y;
                }
            }
            if (MODE > 0) {
                // the scaled panel leaves for the slab: [0] Minv^T, [1..NB] L^T blocks, [NB+1] D^-1 z rows, transposed
                // on the tensor core (I * L^T leaves the transposed block in the C-fragment layout)
                double2 *pan = reinterpret_cast<double2 *>(lws + (size_t)p * LPB);
                __stcs(pan + lane, mit);
#pragma unroll
                for (int b = 0; b < NB1; ++b) {
                    // transpose in the fragment layout: lane (g, t) needs (L[2t][g], L[2t+1][g]), held by lanes
                    // (2t, g >> 1) and (2t + 1, g >> 1) in component g & 1 -- four shuffles instead of two MMAs on
                    // the shared FP64 / tensor pipe
                    const int s0 = 8 * t + (g >> 1), s1 = s0 + 4;
                    const double ax = __shfl_sync(kFull, Ln[b].x, s0), ay = __shfl_sync(kFull, Ln[b].y, s0);
                    const double bx = __shfl_sync(kFull, Ln[b].x, s1), by = __shfl_sync(kFull, Ln[b].y, s1);
                    const bool odd = g & 1;
                    __stcs(pan + (b ? b : NB + 1) * 32 + lane, make_double2(-(odd ? ay : ax), -(odd ? by : bx)));
                }
            }
            WTL(3);
            // ---- trailing update, written one block up and one block left: the window slides with the panel
#pragma unroll
            for (int I = 1; I <= NB; ++I)
#pragma unroll
                for (int J = 1; J <= I; ++J) {
                    double2 c = W[I][J];
                    dmma884(c.x, c.y, Ln[I].x, V[J].x);
                    dmma884(c.x, c.y, Ln[I].y, V[J].y);
                    W[I - 1][J - 1] = c;
                }
#pragma unroll
            for (int J = 1; J <= NB; ++J) {
                double2 c = Rh[J];
                dmma884(c.x, c.y, Ln[0].x, V[J].x);
                dmma884(c.x, c.y, Ln[0].y, V[J].y);
                Rh[J - 1] = c;
            }
            WTL(4);
            // ---- the diagonal block of panel p+1 is factored WHILE block row p+NB+1 is gathered into the staging area:
            //      one instruction stream, two independent dependency chains
            const int q = p + NB1;
            if (p + 1 < NQ) {
                if (q < NQ) {
                    row_prepare(q);
                    __syncwarp();
                    row_gather(q);
                }
                warp_diag_fragment(W[0][0], mi, mit, r2, flagp, lane);
                __syncwarp();
            }
            WTL(2);
            if (q < NQ) {
#pragma unroll
                for (int d = 0; d <= NB; ++d) W[NB][NB - d] = reinterpret_cast<const double2 *>(stage)[d * 32 + lane];
                Rh[NB] = z2;
                if (rowtab[q].y >> 30) Rh[NB] = __ldg(reinterpret_cast<const double2 *>(Q.rhs0 + (size_t)q * 64) + lane);
            } else {
#pragma unroll
                for (int d = 0; d <= NB; ++d) W[NB][d] = z2;
                Rh[NB] = z2;
            }
            WTL(8);
        }

        WTL(1);
        // ---------------- observations: y from the last diagonal block, strains from the accumulated products,
        //                  h = von Mises at the two observed Gauss points (src/fem_postprocess.py:172-185)
        __syncwarp();
        reinterpret_cast<double2 *>(stg)[lane] = mit;  // the last panel's Minv^T [c][k] and 1 / d for the lanes below
        if (g == 0) reinterpret_cast<double2 *>(rd)[t] = r2;
        __syncwarp();
        gacc += __shfl_xor_sync(kFull, gacc, 1);
        gacc += __shfl_xor_sync(kFull, gacc, 2);
        if (t == 0) sG[g] = gacc;
        if (lane < 16) {  // D^-1 L11^-1 e_j for the observed node's dofs j (their unit vectors start in the last panel)
            const int k = lane >> 3, c = lane & 7, j = Q.obs_loc[k];
            nodeL[lane] = (j >= 0) ? stg[j * 8 + c] * rd[c] : 0.0;
        }
        __syncwarp();
        if (lane < 2) {
            const double exx = sG[1 + 3 * lane], eyy = sG[2 + 3 * lane], gxy = sG[3 + 3 * lane];
            double sig[4];
            Tangent C;
            mat_isotropic_plane_strain(mat, exx, eyy, gxy, sig, C);
            double ds[4];
            const double hv = von_mises_ref(sig, ds);
            const double l2m = mat.lam + 2.0 * mat.mu;
            double *o = obs + 8 * lane;
            o[0] = hv;
            o[1] = ds[0] * l2m + ds[1] * mat.lam + ds[2] * mat.lam;  // dh/d(exx)
            o[2] = ds[0] * mat.lam + ds[1] * l2m + ds[2] * mat.lam;  // dh/d(eyy)
            o[3] = ds[3] * mat.mu;                                   // dh/d(gxy)
            o[4] = (ds[0] + ds[1] + ds[2]) * (exx + eyy);            // dh/d(lambda) at fixed u
            o[5] = 2.0 * ds[0] * exx + 2.0 * ds[1] * eyy + ds[3] * gxy;
            if (A.h) A.h[2 * s + lane] = hv;
            // y_k = (L11^-T D^-1 z_f)[j] = sum_c Minv[c][j] lf[c]
            const int j = Q.obs_loc[lane];
            double yv = 0.0;
            if (j >= 0)
                for (int c = 0; c < 8; ++c) yv = fma(stg[j * 8 + c], lf_last[c], yv);
            obs[16 + lane] = yv;
            if (A.y) A.y[2 * s + lane] = yv;
            if (A.f_out) A.f_out[2 * s + lane] = yv;
            if (!(fabs(yv) < 1.0e300) || !(hv < 1.0e300)) *flagp = 1;
        }
        __syncwarp();
        if (MODE > 0) {
            // ---------------- right-hand sides of the reverse pass: v = 0 is u (row 0 = D^-1 z_f); the adjoint
            //                  vectors combine the strain rows and the observed node's unit vectors
            sW[lane] = 0.0;
            sW[32 + lane] = 0.0;
            if (lane < 16) nodew[lane] = 0.0;
            __syncwarp();
            if (lane == 0) {
                sW[0] = 1.0;
                if (MODE == 1) {
                    double gy0, gy1, gh0 = 0.0, gh1 = 0.0;
                    if (A.mode & kElbo) {
                        // d(loss)/d f_j through term2 with the [B, B*S] broadcast (main_custom_training.py:205-214)
                        gy0 = A.gcoef * ((double)A.B * obs[16] - A.ysum[0]);
                        gy1 = A.gcoef * ((double)A.B * obs[17] - A.ysum[1]);
                    } else {
                        gy0 = A.gy[2 * s];
                        gy1 = A.gy[2 * s + 1];
                        gh0 = A.gh[2 * s];
                        gh1 = A.gh[2 * s + 1];
                    }
                    obs[20] = gh0;
                    obs[21] = gh1;
                    for (int i = 0; i < 3; ++i) {
                        sW[8 + 1 + i] = gh0 * obs[1 + i];
                        sW[8 + 4 + i] = gh1 * obs[8 + 1 + i];
                    }
                    nodew[2] = gy0;
                    nodew[3] = gy1;
                } else {
                    // vectors 1, 2: adjoints of y0, y1; 3, 4: adjoints of h0, h1
                    nodew[2 * 1] = 1.0;
                    nodew[2 * 2 + 1] = 1.0;
                    for (int i = 0; i < 3; ++i) {
                        sW[3 * 8 + 1 + i] = obs[1 + i];
                        sW[4 * 8 + 4 + i] = obs[8 + 1 + i];
                    }
                }
            }
            __syncwarp();
            WTL(5);
            // ---------------- reverse pass: x_p = (W Lrhs_p - sum_b x_(p+b) L_(p+b,p)) Minv_p, panels descending,
            //                  fragments straight from the slab (each lane reads back what it stored)
            const double2 Wf = reinterpret_cast<const double2 *>(sW)[lane];
            const double nw0 = nodew[2 * g], nw1 = nodew[2 * g + 1];
            const double2 nl0 = make_double2(nodeL[2 * t], nodeL[2 * t + 1]),
                          nl1 = make_double2(nodeL[8 + 2 * t], nodeL[8 + 2 * t + 1]);
            double2 X[NB1];  // X[b] = -x_(p+b), b = 1..NB
#pragma unroll
            for (int b = 0; b < NB1; ++b) X[b] = z2;
            // u and the adjoint vectors: in the element ring (idle after the forward pass) when they fit, else global
            double *xv = (MODE == 1 && Q.x_in_smem) ? ke : xws;
            double2 cur[NB + 2], nxt[NB + 2], nx2[NB + 2];  // panels p, p-1, p-2: loads two panels ahead of their use
            constexpr int kAhead = 4;  // panels on their way into L2 ahead of the register buffers
            if (lane < (NB + 2) * 4)
                for (int i = 2; i <= 1 + kAhead && NQ - 1 - i >= 0; ++i)
                    prefetch_l2(reinterpret_cast<const char *>(lws + (size_t)(NQ - 1 - i) * LPB) + 128 * lane);
            {
                const double2 *pan = reinterpret_cast<const double2 *>(lws + (size_t)(NQ - 1) * LPB);
#pragma unroll
                for (int b = 0; b < NB + 2; ++b) nxt[b] = __ldcs(pan + b * 32 + lane);
                const double2 *pa2 = reinterpret_cast<const double2 *>(lws + (size_t)(NQ - 2) * LPB);
#pragma unroll
                for (int b = 0; b < NB + 2; ++b) nx2[b] = __ldcs(pa2 + b * 32 + lane);
            }
#pragma unroll 1
            for (int p = NQ - 1; p >= 0; --p) {
#pragma unroll
                for (int b = 0; b < NB + 2; ++b) {
                    cur[b] = nxt[b];
                    nxt[b] = nx2[b];
                }
                if (p > 1) {
                    const double2 *pan = reinterpret_cast<const double2 *>(lws + (size_t)(p - 2) * LPB);
#pragma unroll
                    for (int b = 0; b < NB + 2; ++b) nx2[b] = __ldcs(pan + b * 32 + lane);
                    if (p - 2 - kAhead >= 0 && lane < (NB + 2) * 4)
                        prefetch_l2(reinterpret_cast<const char *>(lws + (size_t)(p - 2 - kAhead) * LPB) + 128 * lane);
                }
                double2 d = z2;
                block_mma<true>(d, Wf, cur[NB + 1], lane);
#pragma unroll
                for (int b = 1; b <= NB; ++b) {
                    double2 c = z2;
                    block_mma<true>(c, X[b], cur[b], lane);
                    d.x += c.x;
                    d.y += c.y;
                }
                if (p == NQ - 1) {
                    d.x += nw0 * nl0.x + nw1 * nl1.x;
                    d.y += nw0 * nl0.y + nw1 * nl1.y;
                }
                double2 x = z2;
                block_mma<true>(x, d, cur[0], lane);
                if (g < NV) *reinterpret_cast<double2 *>(xv + (size_t)g * Q.npad + 8 * p + 2 * t) = x;
#pragma unroll
                for (int b = NB; b > 1; --b) X[b] = X[b - 1];
                X[1] = make_double2(-x.x, -x.y);
            }
            __syncwarp();

            WTL(6);
            // ---------------- element-wise contraction -psi^T (dK/dp) u + explicit dh/dp, chained to x
            constexpr int NADJ = NV - 1;
            double sl[NADJ], sm[NADJ];
#pragma unroll
            for (int v = 0; v < NADJ; ++v) sl[v] = sm[v] = 0.0;
            for (int e = lane; e < Q.nele; e += 32) {  // first-use order: coordinates and band rows are contiguous records
                double xl[4], yl[4], ue[8];
                int lm[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double2 xy = __ldg(reinterpret_cast<const double2 *>(Q.ecoord + (size_t)8 * e) + a);
                    xl[a] = xy.x;
                    yl[a] = xy.y;
                }
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int4 r4 = __ldg(reinterpret_cast<const int4 *>(Q.elm + (size_t)8 * e) + a);
                    lm[4 * a] = r4.x;
                    lm[4 * a + 1] = r4.y;
                    lm[4 * a + 2] = r4.z;
                    lm[4 * a + 3] = r4.w;
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) ue[a] = (lm[a] >= 0) ? xv[lm[a]] : 0.0;
#pragma unroll 1
                for (int gp = 0; gp < 4; ++gp) {
                    ShapeQ4 sh;
                    shapef_q4(xl, yl, gp, M.thk, sh);
                    double uxx, uyy, uxy;
                    strain_q4(sh, ue, uxx, uyy, uxy);
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double *pv = xv + (size_t)(v + 1) * Q.npad;
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
            for (int v = 0; v < NADJ; ++v)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sl[v] += __shfl_down_sync(kFull, sl[v], o);
                    sm[v] += __shfl_down_sync(kFull, sm[v], o);
                }
            if (lane == 0) {
                // d lambda, d mu / d(E, nu), then dE/dx0 = std0 * E ; dnu/dx1 = std1 * nu (1 - 2 nu)
                const double E = E_, nu = nu_;
                const double tt = (1.0 + nu) * (1.0 - 2.0 * nu);
                const double dl_dE = mat.lam / E, dm_dE = mat.mu / E;
                const double dl_dnu = E * (1.0 + 2.0 * nu * nu) / (tt * tt);
                const double dm_dnu = -0.5 * E / ((1.0 + nu) * (1.0 + nu));
                const double dE_dx0 = M.theta_std[0] * E, dnu_dx1 = M.theta_std[1] * nu * (1.0 - 2.0 * nu);
                if (MODE == 1) {
                    const double gh0 = obs[20], gh1 = obs[21];
                    const double gl = -sl[0] + gh0 * obs[4] + gh1 * obs[8 + 4];
                    const double gm = -sm[0] + gh0 * obs[5] + gh1 * obs[8 + 5];
                    A.gx[2 * s] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                    A.gx[2 * s + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                } else {
                    // adjoint vectors v = 0, 1: y0, y1; v = 2, 3: h0, h1 -- the storage order of J
                    double *J = A.ws + (size_t)s * A.ws_stride;
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double gl = -sl[v] + (v >= 2 ? obs[8 * (v - 2) + 4] : 0.0);
                        const double gm = -sm[v] + (v >= 2 ? obs[8 * (v - 2) + 5] : 0.0);
                        J[2 * v] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                        J[2 * v + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && A.status) A.status[s] = *flagp;
        __syncwarp();
        WTL(7);
        WTL_FLUSH;
    }
}

}  // namespace vbfem
