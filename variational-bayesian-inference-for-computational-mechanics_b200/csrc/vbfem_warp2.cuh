// vbfem_warp2.cuh -- the warp-per-sample kernel (vbfem_warp.cuh), second generation: NO per-sample element kernels.
//
// The material of this path is isotropic plane strain with ONE (E, nu) per sample (src/mat_subroutine_tf.py:283-330
// upstream): C_t = lambda m m^T + mu diag(2, 2, 1), hence K(sample) = lambda K_lam + mu K_mu with two
// SAMPLE-INDEPENDENT band matrices.  They are assembled once per mesh (element matrices by the same device routines
// the other kernels use -- shapef_q4 / accumulate_kt with (lambda, mu) = (1, 0) and (0, 1) -- summed on the host in
// element order) and live in shared memory as (K_lam, K_mu) pairs in band form, [row][row - col], 183 KB for Cook
// 20x10, shared by the sixteen warps of the CTA.  Per sample:
//   * the block row entering the register window is eight predicated 16-byte loads and sixteen FMAs per lane --
//     the first generation computed 200 element matrices per sample (Gauss loops, 36 accumulators, a ring of 44
//     matrices in shared memory per warp) and gathered the row through a packed index table;
//   * the adjoint contraction -psi^T (dK/dlambda, dK/dmu) u = -(psi^T K_lam u, psi^T K_mu u) runs INSIDE the
//     reverse pass on the same table: as soon as panel p of u and psi leaves the back substitution it enters a
//     ring of 40 (or hb + 8) rows in shared memory, and block column p + 1 of the band is contracted against the
//     ring in the same basic block as the back substitution of panel p (independent instruction streams): each
//     lane owns a column and every fourth offset below the diagonal, one 16-byte table load and one 16-byte ring
//     load per entry, branch-free.  No solution vector is ever stored, no shape function re-evaluated.
// Without the element kernels the kernel fits 128 registers without spills: sixteen warps per SM in every mode.
// Everything else -- the 8x8 blocked LDL^T in registers, FP64 tensor-core MMAs, the factor slab, the observation
// trick -- is the first generation's.  Results differ from it by rounding only (1e-13 relative).
#pragma once
#include "vbfem_warp.cuh"

namespace vbfem {

// Diagonal block: column by column.  The 2x2-pivot form (warp_diag_fragment_pairs, -DVBFEM_DIAG_PAIRS) halves the
// dependency chain but issues more FP64 instructions: +5 % forward with twelve warps per SM, -3 % with sixteen
// (measured, profiles/README.md) -- with sixteen warps the FP64 pipe, not the chain, is the limit.
#ifdef VBFEM_DIAG_PAIRS
#define WARP2_DIAG warp_diag_fragment_pairs
#else
#define WARP2_DIAG warp_diag_fragment
#endif
// sign flip on the integer pipe (the compiler's own -x is a DADD on the FP64 pipe, the kernel's bottleneck)
__device__ __forceinline__ double neg_alu(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
// A/B build (-DVBFEM_REV_DFMA): the fused mode's back substitution as per-lane FMAs on its two useful vectors instead of
// 8-row MMA fragments.  A quarter of the pipe time, but a longer dependency chain from panel to panel (shuffles, a
// shared-memory round trip for the older panels): 9.3 against 10.8 M fwd+adjoint solves/s -- the reverse pass is bound
// by its chain, not by the pipe.  The tensor-core form is the default.
#ifdef VBFEM_REV_DFMA
constexpr bool kWarp2RevDfma = true;
#else
constexpr bool kWarp2RevDfma = false;
#endif
constexpr int kWarp2Fixed = 640;          // Minv^T and 1/d of the last panel, flag
constexpr int kWarp2Small = 1088;         // small vectors of the observation / reverse pass (136 doubles)
constexpr int kWarp2WinRows = 40;         // ring of u / psi rows: five panels; the contraction of one block column touches
constexpr int kWarp2WinRowsMin = 33;      // hb + 8 <= 33 rows, which is the smallest ring that works (WarpModel::win_rows)
__host__ __device__ constexpr int warp2_smem_per_warp(int nv, int rows = kWarp2WinRows) {
    return (kWarp2Fixed + kWarp2Small + rows * nv * 8 + 15) & ~15;
}

// Element matrices for unit Lame parameters: out[k][0][36] = K_lam of element k, out[k][1][36] = K_mu (lower triangle,
// tri(a, q)); elements in the order of `ecoord`.  Runs once per mesh.
__global__ void warp2_unit_element_kernel(const double *__restrict__ ecoord, int nele, double thk, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nele) return;
    double xl[4], yl[4];
    for (int a = 0; a < 4; ++a) {
        xl[a] = ecoord[8 * k + 2 * a];
        yl[a] = ecoord[8 * k + 2 * a + 1];
    }
    for (int which = 0; which < 2; ++which) {
        Lame mat;
        mat.lam = which ? 0.0 : 1.0;
        mat.mu = which ? 1.0 : 0.0;
        mat.szz = mat.lam;
        double kev[36];
        for (int q = 0; q < 36; ++q) kev[q] = 0.0;
        for (int gp = 0; gp < 4; ++gp) {
            ShapeQ4 sh;
            shapef_q4(xl, yl, gp, thk, sh);
            double sig[4];
            Tangent C;
            mat_isotropic_plane_strain(mat, 0.0, 0.0, 0.0, sig, C);
            accumulate_kt(sh, C, kev);
        }
        for (int q = 0; q < 36; ++q) out[(size_t)k * 72 + which * 36 + q] = kev[q];
    }
}

// MODE 0: y, h   MODE 1: y, h, gx = J^T (gy, gh)   MODE 2: y, h, J = d(y, h)/dx
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) fem_warp2_kernel(const __grid_constant__ DevModel M,
                                                               const __grid_constant__ WarpModel Q,
                                                               const __grid_constant__ Args A) {
    constexpr int NB = kWarpNB, NB1 = NB + 1, LPB = (NB + 2) * 64;
    constexpr int NV = (MODE == 2) ? 5 : 2;
    constexpr int NADJ = NV - 1;
    // Skipping the all-zero right-hand-side work above WarpModel::rhs_first pays with up to three warps per scheduler
    // (+6 % forward, +4 % fused with twelve warps) and in Jacobian mode (+2 %); with four warps per scheduler in
    // forward / fused mode the skipped MMAs were hidden anyway and the extra branches cost 1 %: compiled out there.
    constexpr bool kSkipRhs = NW < 16 || MODE == 2;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ int next_i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const double2 *ktab = reinterpret_cast<const double2 *>(smraw);  // (K_lam, K_mu)[row][row - col], row stride Q.ldt
    const int ldt = Q.ldt;
    unsigned char *wsm = smraw + Q.tab_bytes + (size_t)warp * Q.warp_smem;
    double *stg = reinterpret_cast<double *>(wsm), *rd = stg + 64;  // the last panel's Minv^T and 1 / d (observations)
    int *flagp = reinterpret_cast<int *>(wsm + 576);
    double *small = reinterpret_cast<double *>(wsm + kWarp2Fixed);
    double *sW = small, *nodew = small + 64, *nodeL = small + 80, *sG = small + 96, *lf_last = small + 104,
           *obs = small + 112;
    double *win = reinterpret_cast<double *>(wsm + kWarp2Fixed + kWarp2Small);  // [WR][NV]: u and the adjoint vectors
    const int NQ = Q.NQ, WR = Q.win_rows;
    const int wid = blockIdx.x * NW + warp;
    double *lws = Q.lws + (size_t)wid * Q.lws_stride;
    const double2 z2 = make_double2(0.0, 0.0);
    if (threadIdx.x == 0) next_i = 0;
    {
        const int nvec = Q.tab_bytes / 16;
        const double2 *src = Q.ktab;
        double2 *dst = reinterpret_cast<double2 *>(smraw);
        for (int i = threadIdx.x; i < nvec; i += NW * 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    // Entry (r, c) of the lower band sits at ktab[r * ldt + r - c].  Lane (g, t) of block (q, q - d) holds
    // (r, c) = (8 q + g, 8 (q - d) + 2 t + {0, 1}): the in-row offsets 8 d + g - 2 t - {0, 1} do not depend on q.
    // Outside [0, hb] (and above the diagonal of the diagonal block) the entry is zero.
    int koff[NB1];
    unsigned kmask = 0;
#pragma unroll
    for (int d = 0; d < NB1; ++d) {
        const int o0 = 8 * d + g - 2 * t;  // component x; component y is o0 - 1
        koff[d] = o0;
        if (o0 >= 0 && o0 <= Q.hb) kmask |= 1u << (2 * d);
        if (o0 - 1 >= 0 && o0 - 1 <= Q.hb) kmask |= 2u << (2 * d);
    }

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
        const double lam = mat.lam, mu = mat.mu;
        if (lane == 0) *flagp = 0;
        WTL_DECL;

        // block (q, q - d) of K = lambda K_lam + mu K_mu in the fragment layout
        auto kblock = [&](int q, int d) -> double2 {
            // predicated loads: a branch-free form (clamped address, select) measured 12 % slower in forward mode
            const double2 *row = ktab + (size_t)(8 * q + g) * ldt + koff[d];
            double2 r = z2;
            if (kmask & (1u << (2 * d))) {
                const double2 a = row[0];
                r.x = fma(lam, a.x, mu * a.y);
            }
            if (kmask & (2u << (2 * d))) {
                const double2 a = row[-1];
                r.y = fma(lam, a.x, mu * a.y);
            }
            return r;
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
#pragma unroll
            for (int d = 0; d <= q; ++d) W[q][q - d] = kblock(q, d);
            if ((Q.rhsmask[q >> 6] >> (q & 63)) & 1) Rh[q] = __ldg(reinterpret_cast<const double2 *>(Q.rhs0 + (size_t)q * 64) + lane);
        }

        WTL(0);
        // ---------------- panels
        double gacc = 0.0;  // partial sum over this lane's columns of G[g] = q_g^T K^-1 f
        double2 mi, mit, r2;  // of the current panel: Minv[g][2t..2t+1], its transpose, 1 / d of columns 2t, 2t+1
        WARP2_DIAG(W[0][0], mi, mit, r2, flagp, lane);  // diagonal block of panel 0
#pragma unroll 1
        for (int p = 0; p < NQ; ++p) {
            WTL(1);
            // ---- solve: V = X L11^-T for the right-hand sides (b = 0) and the blocks below; Ln = -V D^-1
            // The right-hand-side rows (load vector, strain functionals) are zero up to their first non-zero block row and
            // stay zero under the elimination until then: for p < rhs_first their solve, their three updates, their slab
            // block and (reverse pass) their product are skipped -- on Cook 20x10 with the reference's observation set-up
            // that is 8 of the 26 forward MMAs for more than half of the panels.
            const bool rhs_live = !kSkipRhs || p >= Q.rhs_first;
            double2 V[NB1], Ln[NB1], Lp[NB1];
            V[0] = Lp[0] = Ln[0] = z2;
#pragma unroll
            for (int b = 0; b < NB1; ++b) {
                if (b == 0 && !rhs_live) continue;
                const double2 xv = b ? W[b][0] : Rh[0];
                double2 v = z2;
                block_mma<true>(v, xv, mi, lane);
                V[b] = v;
                Lp[b] = make_double2(v.x * r2.x, v.y * r2.y);
                Ln[b] = make_double2(neg_alu(Lp[b].x), neg_alu(Lp[b].y));
            }
            if (rhs_live) {  // strain rows against the load row: G[g] += sum_c V[g][c] L[0][c]
                const double lfx = __shfl_sync(kFull, Lp[0].x, t), lfy = __shfl_sync(kFull, Lp[0].y, t);
                gacc = fma(V[0].x, lfx, fma(V[0].y, lfy, gacc));
                if (p == NQ - 1 && g == 0) {
                    lf_last[2 * t] = Lp[0].x;
                    lf_last[2 * t + 1] = Lp[0].y;
                }
            }
            if (MODE > 0) {
                // the scaled panel leaves for the slab: [0] Minv^T, [1..NB] L^T blocks, [NB+1] D^-1 z rows, transposed
                // in the fragment layout by four shuffles per block
                double2 *pan = reinterpret_cast<double2 *>(lws + (size_t)p * LPB);
                __stcs(pan + lane, mit);
#pragma unroll
                for (int b = 0; b < NB1; ++b) {
                    if (b == 0 && !rhs_live) continue;
                    const int s0 = 8 * t + (g >> 1), s1 = s0 + 4;
                    const double ax = __shfl_sync(kFull, Lp[b].x, s0), ay = __shfl_sync(kFull, Lp[b].y, s0);
                    const double bx = __shfl_sync(kFull, Lp[b].x, s1), by = __shfl_sync(kFull, Lp[b].y, s1);
                    const bool odd = g & 1;
                    __stcs(pan + (b ? b : NB + 1) * 32 + lane, make_double2(odd ? ay : ax, odd ? by : bx));
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
                if (rhs_live) {
                    dmma884(c.x, c.y, Ln[0].x, V[J].x);
                    dmma884(c.x, c.y, Ln[0].y, V[J].y);
                }
                Rh[J - 1] = c;
            }
            WTL(4);
            // ---- block row p+NB+1 enters the window (table loads, independent of everything in flight) while the
            //      diagonal block of panel p+1 is factored: one instruction stream, two independent chains
            const int q = p + NB1;
            if (q < NQ) {
#pragma unroll
                for (int d = 0; d <= NB; ++d) W[NB][NB - d] = kblock(q, d);
                Rh[NB] = z2;
                if ((Q.rhsmask[q >> 6] >> (q & 63)) & 1) Rh[NB] = __ldg(reinterpret_cast<const double2 *>(Q.rhs0 + (size_t)q * 64) + lane);
            } else {
#pragma unroll
                for (int d = 0; d <= NB; ++d) W[NB][d] = z2;
                Rh[NB] = z2;
            }
            if (p + 1 < NQ) WARP2_DIAG(W[0][0], mi, mit, r2, flagp, lane);
            WTL(2);
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
            for (int i = lane; i < WR * NV; i += 32) win[i] = 0.0;
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
            double sl[NADJ], sm[NADJ];  // psi_v^T K_lam u, psi_v^T K_mu u: this lane's share
#pragma unroll
            for (int v = 0; v < NADJ; ++v) sl[v] = sm[v] = 0.0;
            // Contraction of block column p of the band: lane (g, t) owns local column cc (g with its two low bits
            // swapped: the (K_lam, K_mu) pairs of a quarter-warp then fall into eight different 16-byte bank groups) and
            // the offsets o = t + 4 j below the diagonal, rows 8 p + cc + o.
            const int cc = (g & 4) | ((g & 1) << 1) | ((g >> 1) & 1);
            constexpr int kTrips = 8;  // offsets up to 31 = the widest band three block sub-diagonals can hold
            int wbase = (8 * (NQ - 1)) % WR;  // window slot of row 8 p
            auto contract = [&](int pc, int wb) {
                int sc = wb + cc;
                sc -= (sc >= WR) ? WR : 0;
                const double *xc = win + sc * NV;
                double uc = xc[0], pcv[NADJ];
#pragma unroll
                for (int v = 0; v < NADJ; ++v) pcv[v] = xc[1 + v];
                const double2 *kcol = ktab + (size_t)(8 * pc + cc + t) * ldt + t;
#pragma unroll
                for (int j = 0; j < kTrips; ++j) {
                    // branch-free: an entry outside the band (or below the last row) reads entry 0 of the table and
                    // counts with a zero coefficient, so that all loads of the column are in flight together
                    if (4 * j > Q.hb) break;  // uniform: the whole trip lies outside the band
                    const int o = t + 4 * j;
                    const bool valid = o <= Q.hb && 8 * pc + cc + o < Q.npad;
                    const double2 kr = valid ? kcol[(size_t)4 * j * (ldt + 1)] : ktab[0];
                    const double kx = valid ? kr.x : 0.0, ky = valid ? kr.y : 0.0;
                    int sr = sc + o;
                    sr -= (sr >= WR) ? WR : 0;
                    const double *xr = win + sr * NV;
                    const double ur = xr[0];
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double pr = xr[1 + v];
                        double sv = pr * ur;                                          // diagonal entry: psi_r K_rr u_r
                        if (j > 0) sv = fma(pr, uc, pcv[v] * ur);                     // o >= 4 > 0
                        else sv = (t == 0) ? sv : fma(pr, uc, pcv[v] * ur);
                        sl[v] = fma(kx, sv, sl[v]);
                        sm[v] = fma(ky, sv, sm[v]);
                    }
                }
            };
            constexpr int kAhead = 4;  // panels on their way into L2 ahead of the register buffers
            if (lane < (NB + 2) * 4)
                for (int i = 2; i <= 1 + kAhead && NQ - 1 - i >= 0; ++i)
                    prefetch_l2(reinterpret_cast<const char *>(lws + (size_t)(NQ - 1 - i) * LPB) + 128 * lane);
            if constexpr (MODE == 1 && kWarp2RevDfma) {
                // ---- Fused mode carries TWO vectors (u, psi): as 8-row MMA fragments six of eight rows would be zeros
                // (10 DMMAs per panel on the pipe that bounds the kernel).  Here lane (v, c, h) = (lane >> 4,
                // (lane >> 1) & 7, lane & 1) owns entry c of vector v and half h of every contraction index: the stored
                // blocks are row-major [c][k] (L^T, Minv^T, the transposed D^-1 z rows), so a lane's four coefficients of
                // a block are 32 contiguous bytes of the slab; 16 + 4 FMAs per lane and panel, two xor-shuffle sums.
                const int v = lane >> 4, c = (lane >> 1) & 7, h = lane & 1;
                double wf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) wf[j] = sW[v * 8 + 4 * h + j];
                const double nwa = nodew[2 * v], nwb = nodew[2 * v + 1], nla = nodeL[c], nlb = nodeL[8 + c];
                double2 ca[NB + 2], cb[NB + 2], na[NB + 2], nb[NB + 2];  // panels p and p-1: k = 4h, 4h+1 | 4h+2, 4h+3
                const int lo2 = c * 4 + 2 * h;                            // double2 index of S[c][4h] inside a block
                {
                    const double2 *pan = reinterpret_cast<const double2 *>(lws + (size_t)(NQ - 1) * LPB) + lo2;
#pragma unroll
                    for (int bk = 0; bk < NB + 2; ++bk) {
                        na[bk] = __ldcs(pan + bk * 32);
                        nb[bk] = __ldcs(pan + bk * 32 + 1);
                    }
                }
                double xprev = 0.0;  // x_(p+1)[v][c] (both halves hold it)
#pragma unroll 1
                for (int p = NQ - 1; p >= 0; --p) {
#pragma unroll
                    for (int bk = 0; bk < NB + 2; ++bk) {
                        ca[bk] = na[bk];
                        cb[bk] = nb[bk];
                    }
                    if (p > 0) {
                        const double2 *pan = reinterpret_cast<const double2 *>(lws + (size_t)(p - 1) * LPB) + lo2;
#pragma unroll
                        for (int bk = 0; bk < NB + 2; ++bk) {
                            na[bk] = __ldcs(pan + bk * 32);
                            nb[bk] = __ldcs(pan + bk * 32 + 1);
                        }
                        if (p - 1 - kAhead >= 0 && lane < (NB + 2) * 4)
                            prefetch_l2(reinterpret_cast<const char *>(lws + (size_t)(p - 1 - kAhead) * LPB) + 128 * lane);
                    }
                    // right-hand-side rows, then the older panels from the ring (their rows 8 (p+b) + 4h + j), the newest
                    // panel last and straight from the registers of its owners
                    double d0 = 0.0, d1 = 0.0;
                    if (!kSkipRhs || p >= Q.rhs_first) {   // no z rows stored below rhs_first
                        d0 = fma(wf[2], cb[NB + 1].x, wf[0] * ca[NB + 1].x);
                        d1 = fma(wf[3], cb[NB + 1].y, wf[1] * ca[NB + 1].y);
                    }
#pragma unroll
                    for (int bk = NB; bk >= 2; --bk) {
                        int r = wbase + 8 * bk + 4 * h;
                        r -= (r >= WR) ? WR : 0;   // WR >= 33: the four rows r .. r+3 may still wrap
                        double xr[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            int rj = r + j;
                            rj -= (rj >= WR) ? WR : 0;
                            xr[j] = (p + bk < NQ) ? win[rj * NV + v] : 0.0;
                        }
                        d0 = fma(-xr[0], ca[bk].x, d0);
                        d1 = fma(-xr[1], ca[bk].y, d1);
                        d0 = fma(-xr[2], cb[bk].x, d0);
                        d1 = fma(-xr[3], cb[bk].y, d1);
                    }
                    {
                        const int src = (v << 4) + 8 * h;  // lane of (v, 4h + j, 0) is src + 2 j
                        const double x0 = __shfl_sync(kFull, xprev, src), x1 = __shfl_sync(kFull, xprev, src + 2);
                        const double x2 = __shfl_sync(kFull, xprev, src + 4), x3 = __shfl_sync(kFull, xprev, src + 6);
                        d0 = fma(-x0, ca[1].x, d0);
                        d1 = fma(-x1, ca[1].y, d1);
                        d0 = fma(-x2, cb[1].x, d0);
                        d1 = fma(-x3, cb[1].y, d1);
                    }
                    double d = d0 + d1;
                    d += __shfl_xor_sync(kFull, d, 1);
                    if (p == NQ - 1) d += nwa * nla + nwb * nlb;
                    double x;
                    {
                        const int src = (v << 4) + 8 * h;
                        const double e0 = __shfl_sync(kFull, d, src), e1 = __shfl_sync(kFull, d, src + 2);
                        const double e2 = __shfl_sync(kFull, d, src + 4), e3 = __shfl_sync(kFull, d, src + 6);
                        const double y0 = fma(e2, cb[0].x, e0 * ca[0].x), y1 = fma(e3, cb[0].y, e1 * ca[0].y);
                        x = y0 + y1;
                        x += __shfl_xor_sync(kFull, x, 1);
                    }
                    xprev = x;
                    WTL(9);
                    if (p + 1 < NQ) contract(p + 1, wbase + 8 >= WR ? wbase + 8 - WR : wbase + 8);
                    __syncwarp();
                    if (h == 0) {
                        int r0 = wbase + c;
                        r0 -= (r0 >= WR) ? WR : 0;
                        win[r0 * NV + v] = x;
                    }
                    __syncwarp();
                    wbase -= 8;
                    wbase += (wbase < 0) ? WR : 0;
                    WTL(10);
                }
            } else {
                double2 cur[NB + 2], nxt[NB + 2], nx2[NB + 2];  // panels p, p-1, p-2: loads two panels ahead of their use
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
                        for (int b = 0; b < NB + 2; ++b)
                            if (b <= NB || !kSkipRhs || p - 2 >= Q.rhs_first) nx2[b] = __ldcs(pan + b * 32 + lane);  // no z rows stored below rhs_first
                        if (p - 2 - kAhead >= 0 && lane < (NB + 2) * 4)
                            prefetch_l2(reinterpret_cast<const char *>(lws + (size_t)(p - 2 - kAhead) * LPB) + 128 * lane);
                    }
                    // one accumulator; the block that needs the newest x (b = 1) comes last: the chain from panel p+1 to
                    // panel p is two block products, the others run ahead
                    double2 d = z2;
                    if (p == NQ - 1) d = make_double2(nw0 * nl0.x + nw1 * nl1.x, nw0 * nl0.y + nw1 * nl1.y);
                    if (!kSkipRhs || p >= Q.rhs_first) block_mma<true>(d, Wf, cur[NB + 1], lane);
    #pragma unroll
                    for (int b = NB; b >= 1; --b) block_mma<true>(d, X[b], cur[b], lane);
                    double2 x = z2;
                    block_mma<true>(x, d, cur[0], lane);
    #pragma unroll
                    for (int b = NB; b > 1; --b) X[b] = X[b - 1];
                    X[1] = make_double2(neg_alu(x.x), neg_alu(x.y));
                    WTL(9);
                    // ---- block column p + 1 of the band is contracted (its rows 8 (p+1) .. 8 (p+1) + hb + 7 are in the window):
                    //      independent of the back-substitution chain above, one basic block with it
                    if (p + 1 < NQ) contract(p + 1, wbase + 8 >= WR ? wbase + 8 - WR : wbase + 8);
                    // ---- panel p of u and the adjoint vectors enters the window (the slot of panel p + 5)
                    __syncwarp();
                    if (g < NV) {
                        int r0 = wbase + 2 * t, r1 = r0 + 1;
                        r0 -= (r0 >= WR) ? WR : 0;
                        r1 -= (r1 >= WR) ? WR : 0;
                        win[r0 * NV + g] = x.x;
                        win[r1 * NV + g] = x.y;
                    }
                    __syncwarp();
                    wbase -= 8;
                    wbase += (wbase < 0) ? WR : 0;
                    WTL(10);
                }
            }
            contract(0, wbase + 8 >= WR ? wbase + 8 - WR : wbase + 8);
            __syncwarp();

            WTL(6);
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
