// vbfem_panel2.cuh -- second generation of the blocked panel kernel for Cook 80x40 (block half bandwidth NB = 11):
// the same mathematics, warp roles, element ring, row records and reverse pass as vbfem_panel.cuh, but the
// off-diagonal TRAILING WINDOW AND THE RIGHT-HAND SIDES LIVE IN THE REGISTERS of the six update warps.
//
// Every update warp owns two whole block diagonals of the window (the right-hand-side row counts as one): the
// live blocks of diagonal d at panel p are (p+J+d, p+J), J = 0..NB-d -- NB+1-d of them -- and the pairs
// (rhs 12 + d11 1), (d1 11 + d10 2), (d2 10 + d9 3), (d3 9 + d8 4), (d4 8 + d7 5), (d5 7 + d6 6) give every warp
// exactly 13 blocks (26 doubles per lane), two blocks to solve and eleven to update per panel.  Position J of a
// diagonal is a fixed register: J = 0 is solved (V = X L11^-T published to shared memory for everybody's B
// fragments, -V D^-1 for the A fragments), J >= 1 takes C -= L V^T and is written to position J-1 (the MMA's D
// operand: the window slides for free, as in vbfem_warp.cuh), the last position receives the block of the entering
// row.  The update is two fragment loads and two MMAs per block with compile-time roles: no accumulator load, no
// store, no address table, no load-to-use chain through shared memory (the first generation spent 47 % of its
// stall samples there).  Only the diagonal ring stays in shared memory (the look-ahead warp factors from it).
// The three warp roles run their own copy of the panel loop and meet at two named barriers per panel.
#pragma once
#include "vbfem_panel.cuh"

namespace vbfem {

constexpr int kPanel2NB = 11;  // the register layout above is laid out for this block half bandwidth

template <int V>
struct IC {
    static constexpr int value = V;
};

// MODE 0: y, h   MODE 1: y, h, gx = J^T (gy, gh)   MODE 2: y, h, J = d(y, h)/dx
template <int MODE, bool DMMA>
__global__ void __launch_bounds__(kPanelNT, 2) fem_panel2_kernel(const __grid_constant__ DevModel M,
                                                                const __grid_constant__ PanelModel Q,
                                                                const __grid_constant__ Args A) {
    extern __shared__ __align__(16) unsigned char smraw[];
    PanelSmem &S = *reinterpret_cast<PanelSmem *>(smraw);
    double *wdiag = reinterpret_cast<double *>(smraw + Q.o_wdiag);  // ring of NB+2 diagonal blocks (block (c, c): slot c mod (NB+2))
    double *fresh = reinterpret_cast<double *>(smraw + Q.o_fresh);  // the entering block row by diagonal d = 0..NB, then its rhs block
    double *vst = reinterpret_cast<double *>(smraw + Q.o_vst);      // V blocks of the current panel (block row p+J at J), C layout
    double *lst = reinterpret_cast<double *>(smraw + Q.o_lst);  // two staging panels (transposed, scaled)
    double *ke = reinterpret_cast<double *>(smraw + Q.o_ke);    // R element matrices (36 each), then 0.0, 1.0
    double *lneg = reinterpret_cast<double *>(smraw + Q.o_lneg);  // -L blocks of the current panel, then a dummy block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NB = Q.NB, NQ = Q.NQ, NB1 = NB + 1, NB2 = NB + 2, LPB = (NB + 2) * 64;  // LPB: doubles per stored panel
    constexpr int NV = (MODE == 2) ? 5 : 2;
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

        // ---------------- reset: diagonal ring to zero, ring constants, slot tables
        {
            double2 *w2 = reinterpret_cast<double2 *>(wdiag);
            for (int i = tid; i < NB2 * 32; i += kPanelNT) w2[i] = make_double2(0.0, 0.0);
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
#ifdef VBFEM_PANEL_DIAG_PAIRS
            // 2x2 pivot blocks: columns (k, k+1) are eliminated together.  With A = a[k][k], B = a[k+1][k], C = a[k+1][k+1],
            // det = A C - B^2 and p = a[.][k], q = a[.][k+1], the trailing entries take
            //     a[i][j] -= (p_i (C p_j - B q_j) + q_i (A q_j - B p_j)) / det
            // with everything but 1 / det computed while the reciprocal is in flight: the chain of a PAIR of pivots is
            // det (2 FP64) -> reciprocal -> one FMA, against two times reciprocal -> FMA.  A/B build only: measured 190 / 122 k
            // against 197 / 126 k solves/s (forward / fused) for the column-by-column form below -- 70 % more FP64
            // instructions and, at 128 registers next to the 36-entry working set, 230 bytes of spills.
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double A_ = a[tri(k, k)], B_ = a[tri(k + 1, k)], C_ = a[tri(k + 1, k + 1)];
                const double det = fma(A_, C_, -B_ * B_);
                bad |= (unsigned)(__double2hiint(A_) - 0x00200000) >= 0x7fd00000u;
                bad |= (unsigned)(__double2hiint(det) - 0x00200000) >= 0x7fd00000u;
                const double rdet = fast_rcp3(det);
                const double ra = fast_rcp3(A_);
                rdv[k] = ra;
                rdv[k + 1] = A_ * rdet;
#pragma unroll
                for (int j = k + 2; j < 8; ++j) {
                    const double pj = a[tri(j, k)], qj = a[tri(j, k + 1)];
                    const double ns = fma(C_, pj, -B_ * qj), nt = fma(A_, qj, -B_ * pj);
#pragma unroll
                    for (int i = j; i < 8; ++i) {
                        const double w = fma(a[tri(i, k)], ns, a[tri(i, k + 1)] * nt);
                        a[tri(i, j)] = fma(-w, rdet, a[tri(i, j)]);
                    }
                    a[tri(j, k)] = pj * ra;        // L[j][k]     (p_j, q_j are dead from here on)
                    a[tri(j, k + 1)] = nt * rdet;  // L[j][k+1]
                }
                a[tri(k + 1, k)] = B_ * ra;
            }
#else
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = a[tri(k, k)];
                // positive, finite, not tiny: sign and exponent bits only (integer pipe; the FP64 pipe is the contended one)
                bad |= (unsigned)(__double2hiint(d) - 0x00200000) >= 0x7fd00000u;
                rdv[k] = fast_rcp3(d);
                // the next pivot first, one level after the reciprocal: it heads the dependency chain of the block --
                // and this warp's chain is the critical path of the panel
                if (k + 1 < 8) {
                    const double sq = a[tri(k + 1, k)] * a[tri(k + 1, k)];
                    a[tri(k + 1, k + 1)] = fma(-sq, rdv[k], a[tri(k + 1, k + 1)]);
                }
#pragma unroll
                for (int j = k + 1; j < 8; ++j) {
                    const double ljk = a[tri(j, k)] * rdv[k];
#pragma unroll
                    for (int i = j; i < 8; ++i)
                        if (!(i == k + 1 && j == k + 1)) a[tri(i, j)] = fma(-a[tri(i, k)], ljk, a[tri(i, j)]);
                    a[tri(j, k)] = ljk;  // rows i > j of column k stay unscaled until their own turn
                }
            }
#endif
            if (bad && lane == 0) S.flag = 1;
            // column j of the inverse of the unit factor, column oriented: dependency depth 7 instead of 28
            const int j = lane & 7;
            double m[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int i = k + 1; i < 8; ++i) m[i] = fma(-a[tri(i, k)], m[k], m[i]);  // m[k] = 0 for k < j
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

        // ---------------- the first NB+1 block rows enter one by one: all warps gather row q into the staging area
        //                  (its diagonal block straight into the diagonal ring), the slot owners pick their blocks up
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
        auto bsync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        const int nrow0 = (NB1 < NQ) ? NB1 : NQ;  // rows that fill the window before the first panel
        // The three warp roles run their own copy of the prologue and of the panel loop (the register blocks of the
        // update warps then never overlap the 36-entry working set of the look-ahead warp's factorisation) and meet
        // at named barriers: three per prologue row, two per panel.
        if (warp < kPanelUpdW) {
            auto update_role = [&](auto na_, auto nb_, auto da_, auto db_) {
                constexpr int NA_ = decltype(na_)::value, NB_ = decltype(nb_)::value;  // live blocks of the two diagonals
                constexpr int dA = decltype(da_)::value, dB = decltype(db_)::value;    // kPanel2NB + 1: the right-hand-side row
                constexpr int NB = kPanel2NB, NB1 = NB + 1, NB2 = NB + 2;               // compile-time here: static block roles
                const double2 *fA = reinterpret_cast<const double2 *>(fresh + dA * 64) + lane;
                const double2 *fB = reinterpret_cast<const double2 *>(fresh + dB * 64) + lane;
                double2 CA[NA_], CB[NB_];
#pragma unroll
                for (int J = 0; J < NA_; ++J) CA[J] = make_double2(0.0, 0.0);
#pragma unroll
                for (int J = 0; J < NB_; ++J) CB[J] = make_double2(0.0, 0.0);
                // ---- prologue: block row q holds block (q, q-d) of diagonal d = position q-d; rhs block q = position q
#pragma unroll
                for (int q = 0; q < NB1; ++q) {  // the host plan guarantees NQ > NB + 1
                    bsync();
                    bsync();
#pragma unroll
                    for (int J = 0; J < NA_; ++J)
                        if (J == (dA <= NB ? q - dA : q)) CA[J] = *fA;
#pragma unroll
                    for (int J = 0; J < NB_; ++J)
                        if (J == q - dB) CB[J] = *fB;
                    bsync();
                }
                bsync();
                bsync();
                PTL(0);
                // ---- panels
                int rslot = 0;  // p mod (NB+2): slot of panel p in the diagonal ring
                for (int p = 0; p < NQ; ++p) {
                    const int par = p & 1;
                    double *stg = lst + par * LPB;  // staging panel: [0] inverse unit factor^T, [1..NB] L^T blocks, [NB+1] rhs
                    const double2 r2 = reinterpret_cast<const double2 *>(S.rd[par])[t];
                    const double2 mi = reinterpret_cast<const double2 *>(S.minv[par])[lane];  // Minv[g][2t..2t+1]
                    // phase B: the blocks of the entering row move in; position 0 of both diagonals is in the panel
                    // column: V = X L11^-T, published with -V D^-1 and (adjoint) the scaled transpose
                    if (p > 0) {
                        CA[NA_ - 1] = *fA;
                        CB[NB_ - 1] = *fB;
                    }
                    double2 vA = make_double2(0.0, 0.0), vB = vA;
                    if (DMMA) {  // four independent MMAs, then two adds: no MMA waits for another one
                        double2 wA = vA, wB = vB;
                        dmma884(vA.x, vA.y, CA[0].x, mi.x);
                        dmma884(vB.x, vB.y, CB[0].x, mi.x);
                        dmma884(wA.x, wA.y, CA[0].y, mi.y);
                        dmma884(wB.x, wB.y, CB[0].y, mi.y);
                        vA = make_double2(vA.x + wA.x, vA.y + wA.y);
                        vB = make_double2(vB.x + wB.x, vB.y + wB.y);
                    } else {
                        block_mma<false>(vA, CA[0], mi, lane);
                        block_mma<false>(vB, CB[0], mi, lane);
                    }
                    const double2 lA = make_double2(vA.x * r2.x, vA.y * r2.y), lB = make_double2(vB.x * r2.x, vB.y * r2.y);
                    reinterpret_cast<double2 *>(lneg + dA * 64)[lane] = make_double2(-lA.x, -lA.y);
                    reinterpret_cast<double2 *>(lneg + dB * 64)[lane] = make_double2(-lB.x, -lB.y);
                    if (dA <= NB) reinterpret_cast<double2 *>(vst + dA * 64)[lane] = vA;
                    reinterpret_cast<double2 *>(vst + dB * 64)[lane] = vB;
                    if (MODE > 0) {
                        // the stored panel holds L^T blocks: transposed in the fragment layout by four shuffles
                        // (lane (g, t) needs L[2t][g], L[2t+1][g], held by lanes (2t, g >> 1), (2t+1, g >> 1) in
                        // component g & 1)
                        const int s0 = 8 * t + (g >> 1), s1 = s0 + 4;
                        const bool odd = g & 1;
                        {
                            const double ax = __shfl_sync(kFull, lA.x, s0), ay = __shfl_sync(kFull, lA.y, s0);
                            const double bx = __shfl_sync(kFull, lA.x, s1), by = __shfl_sync(kFull, lA.y, s1);
                            reinterpret_cast<double2 *>(stg + dA * 64)[lane] = make_double2(odd ? ay : ax, odd ? by : bx);
                        }
                        {
                            const double ax = __shfl_sync(kFull, lB.x, s0), ay = __shfl_sync(kFull, lB.y, s0);
                            const double bx = __shfl_sync(kFull, lB.x, s1), by = __shfl_sync(kFull, lB.y, s1);
                            reinterpret_cast<double2 *>(stg + dB * 64)[lane] = make_double2(odd ? ay : ax, odd ? by : bx);
                        }
                        fence_async_smem();
                    }
                    if (dA > NB) {  // warp 0: strain rows against the load row, G[g] += sum_c V[g][c] L[0][c]
                        const double lfx = __shfl_sync(kFull, lA.x, t), lfy = __shfl_sync(kFull, lA.y, t);
                        gacc = fma(vA.x, lfx, fma(vA.y, lfy, gacc));
                        if (p == NQ - 1 && g == 0) {
                            S.lf_last[2 * t] = lA.x;
                            S.lf_last[2 * t + 1] = lA.y;
                        }
                    }
                    PTL(1);
                    bsync();
                    PTL(2);
                    // phase C: trailing update, written one position down: two fragment loads and two MMAs per block
                    {
                        const double2 *la = reinterpret_cast<const double2 *>(lneg) + lane;
                        const double2 *vb = reinterpret_cast<const double2 *>(vst) + lane;
#pragma unroll
                        for (int J = 1; J < NA_; ++J) {
                            const double2 a = la[(dA <= NB ? J + dA : NB1) * 32], b = vb[J * 32];
                            double2 c = CA[J];
                            if (DMMA) {
                                dmma884(c.x, c.y, a.x, b.x);
                                dmma884(c.x, c.y, a.y, b.y);
                            } else {
                                block_mma<false>(c, a, b, lane);
                            }
                            CA[J - 1] = c;
                        }
#pragma unroll
                        for (int J = 1; J < NB_; ++J) {
                            const double2 a = la[(J + dB) * 32], b = vb[J * 32];
                            double2 c = CB[J];
                            if (DMMA) {
                                dmma884(c.x, c.y, a.x, b.x);
                                dmma884(c.x, c.y, a.y, b.y);
                            } else {
                                block_mma<false>(c, a, b, lane);
                            }
                            CB[J - 1] = c;
                        }
                    }
                    // the diagonal blocks (J, J), J = 2..NB, stay in the shared-memory ring (block (1, 1) belongs to
                    // the look-ahead warp)
                    for (int J = 2 + warp; J <= NB; J += kPanelUpdW) {
                        double2 *D = reinterpret_cast<double2 *>(wdiag + wrap(rslot + J, NB2) * 64);
                        double2 c = D[lane];
                        block_mma<DMMA>(c, reinterpret_cast<const double2 *>(lneg + J * 64)[lane],
                                        reinterpret_cast<const double2 *>(vst + J * 64)[lane], lane);
                        D[lane] = c;
                    }
                    if (MODE > 0 && dA == 5 && lane == 0) {
                        // the finished panel leaves for HBM.  The update warps reach the barrier long before the look-ahead
                        // warp does: one of them waits here until the copy has read the staging buffer, so that the
                        // look-ahead warp (which writes the next panel's Minv^T into the other buffer) never has to
                        bulk_store(lws + (size_t)p * LPB, stg, LPB * 8);
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    PTL(3);
                    bsync();
                    PTL(4);
                    rslot = (rslot + 1 == NB2) ? 0 : rslot + 1;
                }
            };
            switch (warp) {
                case 0: update_role(IC<12>{}, IC<1>{}, IC<12>{}, IC<11>{}); break;  // right-hand sides + diagonal 11
                case 1: update_role(IC<11>{}, IC<2>{}, IC<1>{}, IC<10>{}); break;   // diagonals 1 and 10
                case 2: update_role(IC<10>{}, IC<3>{}, IC<2>{}, IC<9>{}); break;
                case 3: update_role(IC<9>{}, IC<4>{}, IC<3>{}, IC<8>{}); break;
                case 4: update_role(IC<8>{}, IC<5>{}, IC<4>{}, IC<7>{}); break;
                default: update_role(IC<7>{}, IC<6>{}, IC<5>{}, IC<6>{}); break;    // diagonals 5 and 6
            }
        } else {
            // ---- prologue (warps 6 and 7, 64 threads): row q is gathered into the staging area, its diagonal block
            //      straight into the diagonal ring; the update warps pick their blocks up between the barriers
            const int t64 = tid - kPanelUpdW * 32;
            for (int q = 0; q < nrow0; ++q) {
                {
                    double2 *f2 = reinterpret_cast<double2 *>(fresh);
                    for (int i = t64; i < NB2 * 32; i += 64) f2[i] = make_double2(0.0, 0.0);
                }
                bsync();
                const unsigned char *rc = Q.rec + (size_t)q * Q.rec_stride;
                const int4 hd = *reinterpret_cast<const int4 *>(rc);  // new elements, entries, first new element
                if (t64 < 32) reinterpret_cast<double2 *>(fresh + NB1 * 64)[t64] = reinterpret_cast<const double2 *>(rc + 16)[t64];
                const ushort4 *src = reinterpret_cast<const ushort4 *>(rc + Q.rec_o_src);
                const unsigned short *dstp = reinterpret_cast<const unsigned short *>(rc + Q.rec_o_dst);
                for (int i = t64; i < hd.y; i += 64) {
                    const int dst = dstp[i];
                    const ushort4 sr = src[i];
                    const int d = dst >> 6;
                    double *blk = d ? fresh + d * 64 : wdiag + q * 64;  // q <= NB: slot q of the diagonal ring
                    blk[dst & 63] = ((ke[sr.x] + ke[sr.y]) + ke[sr.z]) + ke[sr.w];
                }
                bsync();
                bsync();
            }
            fence_async_smem();
            bsync();
            if (tid == 7 * 32)  // ring slots of elements the first rows no longer need may now be overwritten
                for (int j = 0; j < kPanelRecDepth && j < nrec; ++j) fetch_row(NB1 + j, (int)((rec_base + j) % kPanelRecDepth));
            if (warp == 6) diag_factor(wdiag, lst, S.rd[0], S.minv[0]);  // block (0, 0): slot 0 of the diagonal ring
            bsync();
            PTL(0);
            int rslot = 0;
            if (warp == 6) {
                for (int p = 0; p < NQ; ++p) {
                    const int par = p & 1;
                    double *stg = lst + par * LPB;
                    PTL(1);
                    bsync();
                    PTL(2);
                    // block (p+1, p+1) gets its update ahead of the others and is factored at once, so that the next panel's
                    // solve can start right after the barrier.  This warp is the critical path of the panel: nothing else here.
                    if (p + 1 < NQ) {
                        __syncwarp();
                        double *Dn = wdiag + wrap(rslot + 1, NB2) * 64;
                        double2 c = reinterpret_cast<double2 *>(Dn)[lane];
                        block_mma<DMMA>(c, reinterpret_cast<const double2 *>(lneg + 64)[lane],
                                        reinterpret_cast<const double2 *>(vst + 64)[lane], lane);
                        reinterpret_cast<double2 *>(Dn)[lane] = c;
                        __syncwarp();
                        diag_factor(Dn, lst + (par ^ 1) * LPB, S.rd[par ^ 1], S.minv[par ^ 1]);
                    }
                    PTL(3);
                    bsync();
                    PTL(4);
                    rslot = (rslot + 1 == NB2) ? 0 : rslot + 1;
                }
            } else {
                for (int p = 0; p < NQ; ++p) {
                    PTL(1);
                    bsync();
                    PTL(2);
                    // block row q = p+NB+1 is assembled for the positions that fall free: diagonal block into the spare
                    // slot of the diagonal ring, the others and the right-hand-side block into the staging area -- all
                    // from the row's record, which a bulk copy brought into shared memory several panels ago
                    const int q = p + NB1;
                    const double2 z2 = make_double2(0.0, 0.0);
                    double *Dq = wdiag + wrap(rslot + NB1, NB2) * 64;
                    reinterpret_cast<double2 *>(Dq)[lane] = z2;
                    for (int d = 1; d <= NB1; ++d) reinterpret_cast<double2 *>(fresh + d * 64)[lane] = z2;
                    if (q < NQ) {
                        const unsigned use = rec_base + (unsigned)p;
                        const int sl = (int)(use % kPanelRecDepth);
                        const unsigned char *rc = recs + sl * Q.rec_stride;
                        mbar_wait(&S.rbar[sl], (use / kPanelRecDepth) & 1u);
                        const int4 hd = *reinterpret_cast<const int4 *>(rc);
                        __syncwarp();
                        reinterpret_cast<double2 *>(fresh + NB1 * 64)[lane] = reinterpret_cast<const double2 *>(rc + 16)[lane];
                        const ushort4 *src = reinterpret_cast<const ushort4 *>(rc + Q.rec_o_src);
                        const unsigned short *dstp = reinterpret_cast<const unsigned short *>(rc + Q.rec_o_dst);
                        for (int i = lane; i < hd.y; i += 32) {
                            const int dst = dstp[i];
                            const ushort4 sr = src[i];
                            const int d = dst >> 6;
                            double *blk = d ? fresh + d * 64 : Dq;
                            blk[dst & 63] = ((ke[sr.x] + ke[sr.y]) + ke[sr.z]) + ke[sr.w];
                        }
                        __syncwarp();
                        if (lane == 0 && p + kPanelRecDepth < nrec) fetch_row(q + kPanelRecDepth, sl);  // this slot's next tenant
                    }
                    PTL(3);
                    bsync();
                    PTL(4);
                    rslot = (rslot + 1 == NB2) ? 0 : rslot + 1;
                }
            }
        }
        __syncthreads();

        // ---------------- observations: y from the last diagonal block, strains from the accumulated
        //                  products, h = von Mises at the two observed Gauss points (src/fem_postprocess.py:172-185)
        const double *stgl = lst + ((NQ - 1) & 1) * LPB;  // last panel: [c][k] = Minv[k][c]
        const double *rdl = S.rd[(NQ - 1) & 1];
        if (MODE > 0 && tid == 5 * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
            double *stage0 = reinterpret_cast<double *>(smraw + Q.o_big);  // bulk-load ring over the forward pass's (now idle) areas
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
