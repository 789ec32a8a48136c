// vbfem_panel2.cuh -- second generation of the blocked panel kernel for wide bands (Cook 80x40): the same
// mathematics, warp roles, element ring, row records and reverse pass as vbfem_panel.cuh, but the TRAILING WINDOW
// LIVES IN REGISTERS.  The (NB+1) NB / 2 + ... off-diagonal window blocks and the right-hand-side ring -- 90 ring
// slots for NB = 11 -- are dealt to the six update warps, 15 slots (30 doubles per lane) each.  A slot keeps its
// place in its diagonal's ring (slot r of diagonal d holds the block whose column is congruent to r modulo the
// ring length), so its ROLE changes from panel to panel -- J = (r - p) mod len: 0 = in the panel column (solve,
// publish V and -L to shared memory), 1..len-2 = trailing block (update), len-1 = spare (the entering row's
// block is being assembled for it) -- while its REGISTERS never move.  The update is then two MMAs on two
// fragment loads: no accumulator load, no store, no load-to-use chain through shared memory per block (the first
// generation spent 47 % of its stall samples there).  Only the diagonal ring stays in shared memory (the
// look-ahead warp factors from it).
#pragma once
#include "vbfem_panel.cuh"

namespace vbfem {

constexpr int kPanel2Slots = 15;  // register slots per update warp (6 x 15 = 90 ring slots for NB = 11)

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

        // ---------------- reset: diagonal ring and register slots to zero, ring constants, slot tables
        double2 C[kPanel2Slots];  // this warp's window / right-hand-side ring slots (update warps)
#pragma unroll
        for (int k = 0; k < kPanel2Slots; ++k) C[k] = make_double2(0.0, 0.0);
        {
            double2 *w2 = reinterpret_cast<double2 *>(wdiag);
            for (int i = tid; i < NB2 * 32; i += kPanelNT) w2[i] = make_double2(0.0, 0.0);
            if (tid == 0) {
                ke[Q.R * 36] = 0.0;
                ke[Q.R * 36 + 1] = 1.0;
                S.flag = 0;
            }
            if (tid <= NB2) S.colslot[0][tid] = 0;
            if (tid < 32) reinterpret_cast<double2 *>(lneg + (NB + 2) * 64)[tid] = make_double2(0.0, 0.0);  // the zero block
        }
        __syncthreads();

        double gacc = 0.0;  // update warps: partial sum of G[g] over the columns whose rhs block this warp solved
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
        // ring length of diagonal d (d = NB+1: the right-hand-side ring)
        auto ring_len = [&](int d) { return d <= NB ? NB2 - d : NB2; };
        for (int q = 0; q <= NB && q < NQ; ++q) {
            {
                double2 *f2 = reinterpret_cast<double2 *>(fresh);
                for (int i = tid; i < NB2 * 32; i += kPanelNT) f2[i] = make_double2(0.0, 0.0);
            }
            __syncthreads();
            const unsigned char *rc = Q.rec + (size_t)q * Q.rec_stride;
            const int4 hd = *reinterpret_cast<const int4 *>(rc);  // new elements, entries, first new element
            if (tid < 32) reinterpret_cast<double2 *>(fresh + NB1 * 64)[tid] = reinterpret_cast<const double2 *>(rc + 16)[tid];
            const ushort4 *src = reinterpret_cast<const ushort4 *>(rc + Q.rec_o_src);
            const unsigned short *dstp = reinterpret_cast<const unsigned short *>(rc + Q.rec_o_dst);
            for (int i = tid; i < hd.y; i += kPanelNT) {
                const int dst = dstp[i];
                const ushort4 sr = src[i];
                const int d = dst >> 6;
                double *blk = d ? fresh + d * 64 : wdiag + q * 64;  // q <= NB: slot q of the diagonal ring
                blk[dst & 63] = ((ke[sr.x] + ke[sr.y]) + ke[sr.z]) + ke[sr.w];
            }
            __syncthreads();
            if (warp < kPanelUpdW) {
#pragma unroll
                for (int k = 0; k < kPanel2Slots; ++k) {
                    const int d = Q.slot_d[warp][k], r = Q.slot_r[warp][k];
                    if (d == 0) continue;
                    const int len = ring_len(d);
                    const bool mine = (d <= NB) ? (d <= q && r == (q - d) % len) : (r == q % len);
                    if (mine) C[k] = reinterpret_cast<const double2 *>(fresh + (d <= NB ? d : NB1) * 64)[lane];
                }
            }
            __syncthreads();
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 7 * 32)  // ring slots of elements the first rows no longer need may now be overwritten
            for (int j = 0; j < kPanelRecDepth && j < nrec; ++j) fetch_row(NB1 + j, (int)((rec_base + j) % kPanelRecDepth));
        if (warp == 6) diag_factor(wdiag, lst, S.rd[0], S.minv[0]);  // block (0, 0): slot 0 of the diagonal ring
        __syncthreads();
        PTL(0);

        // ---------------- panels.  The three warp roles run their own copy of the panel loop (the register slots
        //                  of the update warps then never overlap the 36-entry working set of the look-ahead
        //                  warp's factorisation) and meet at two block barriers per panel.
        auto bsync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        if (warp < kPanelUpdW) {
            const double2 idf = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);  // identity fragment
            int rslot = 0;  // p mod (NB+2): slot of panel p in the rhs ring and in the diagonal ring
            // per slot, in one register: J = (r - p) mod len (bits 0..7), len (8..15), d (16..23); J counts down
            int meta[kPanel2Slots];
#pragma unroll
            for (int k = 0; k < kPanel2Slots; ++k) {
                const int d = Q.slot_d[warp][k], r = Q.slot_r[warp][k];
                meta[k] = d ? (r | (ring_len(d) << 8) | (d << 16)) : 0;
            }
            for (int p = 0; p < NQ; ++p) {
                const int par = p & 1;
                double *stg = lst + par * LPB;  // staging panel: [0] inverse unit factor^T, [1..NB] L^T blocks, [NB+1] rhs
                const double2 r2 = reinterpret_cast<const double2 *>(S.rd[par])[t];
                const double2 mi = reinterpret_cast<const double2 *>(S.minv[par])[lane];  // Minv[g][2t..2t+1]
                // ---- phase B: every slot learns its role in this panel, J = (r - p) mod len.  J = len-2: the block
                //      assembled during the previous panel moves in.  J = 0: the block is in the panel column:
                //      V = X L11^-T goes to shared memory for everybody's B fragments, -V D^-1 for the A fragments,
                //      the scaled transpose to the staging panel.
                // pass 1 (straight-line, registers only): roles, the entering blocks, the shared-memory offsets of the
                // update's fragments (a slot that takes no update this panel multiplies by the zero block behind lneg)
                unsigned offs[kPanel2Slots];  // (A offset) | (B offset << 16), in units of 16 bytes from lneg / vst
                unsigned solve_mask = 0;
#pragma unroll
                for (int k = 0; k < kPanel2Slots; ++k) {
                    const int J = meta[k] & 255, len = (meta[k] >> 8) & 255, d = meta[k] >> 16;
                    const bool used = d != 0;
                    if (used && J == len - 2 && p > 0)
                        C[k] = reinterpret_cast<const double2 *>(fresh + (d <= NB ? d : NB1) * 64)[lane];
                    if (used && J == 0) solve_mask |= 1u << k;
                    const bool upd = used && J >= 1 && J <= len - 2;
                    const int I = upd ? ((d <= NB) ? J + d : NB1) : NB + 2;  // NB+2: the zero block
                    offs[k] = (unsigned)(I * 32) | ((unsigned)((upd ? J : 1) * 32) << 16);
                    if (used) meta[k] += J ? -1 : len - 1;  // the slot's J in the next panel
                }
                // pass 2: the (two or three) slots of this warp that sit in the panel column -- one copy of the solve
                while (solve_mask) {
                    const int k = __ffs(solve_mask) - 1;
                    solve_mask &= solve_mask - 1;
                    double2 xv = make_double2(0.0, 0.0);
                    int d = 0;
#pragma unroll
                    for (int kk = 0; kk < kPanel2Slots; ++kk)
                        if (kk == k) {
                            xv = C[kk];
                            d = meta[kk] >> 16;
                        }
                    const int I = (d <= NB) ? d : NB1;  // block row p+d of the window, or the right-hand sides
                    double2 v = make_double2(0.0, 0.0);
                    block_mma<DMMA>(v, xv, mi, lane);
                    const double2 l = make_double2(v.x * r2.x, v.y * r2.y);
                    reinterpret_cast<double2 *>(lneg + I * 64)[lane] = make_double2(-l.x, -l.y);
                    if (d <= NB) reinterpret_cast<double2 *>(vst + I * 64)[lane] = v;
                    if (MODE > 0) {
                        double2 lt = make_double2(0.0, 0.0);
                        block_mma<DMMA>(lt, idf, l, lane);
                        reinterpret_cast<double2 *>(stg + I * 64)[lane] = lt;
                    }
                    if (d > NB) {  // strain rows against the load row, G[g] += sum_c V[g][c] L[0][c]
                        const double lfx = __shfl_sync(kFull, l.x, t), lfy = __shfl_sync(kFull, l.y, t);
                        gacc = fma(v.x, lfx, fma(v.y, lfy, gacc));
                        if (p == NQ - 1 && g == 0) {
                            S.lf_last[2 * t] = l.x;
                            S.lf_last[2 * t + 1] = l.y;
                        }
                    }
                }
                if (MODE > 0) fence_async_smem();
                PTL(1);
                bsync();
                PTL(2);
                // ---- phase C: trailing update C(I,J) -= L_I V_J^T on the register slots: two fragment loads and two
                //      MMAs per slot, no branch -- the loads of the next slots run ahead of the MMAs
                {
                    const double2 *la = reinterpret_cast<const double2 *>(lneg) + lane;
                    const double2 *vb = reinterpret_cast<const double2 *>(vst) + lane;
#pragma unroll
                    for (int k = 0; k < kPanel2Slots; ++k) {
                        const double2 a = la[offs[k] & 0xffffu];
                        const double2 b = vb[offs[k] >> 16];
                        if (DMMA) {
                            dmma884(C[k].x, C[k].y, a.x, b.x);
                            dmma884(C[k].x, C[k].y, a.y, b.y);
                        } else {
                            block_mma<false>(C[k], a, b, lane);
                        }
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
                PTL(3);
                bsync();
                PTL(4);
                rslot = (rslot + 1 == NB2) ? 0 : rslot + 1;
            }
        } else if (warp == 6) {
            int rslot = 0;
            for (int p = 0; p < NQ; ++p) {
                const int par = p & 1;
                double *stg = lst + par * LPB;
                PTL(1);
                bsync();
                PTL(2);
                // the finished panel leaves for HBM; then block (p+1, p+1) gets its update ahead of the others and
                // is factored at once, so that the next panel's solve can start right after the barrier
                if (MODE > 0 && lane == 0) {
                    bulk_store(lws + (size_t)p * LPB, stg, LPB * 8);
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // panel p-1 has left the other buffer
                }
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
            int rslot = 0;
            for (int p = 0; p < NQ; ++p) {
                const int par = p & 1;
                const int *cs = S.colslot[par];
                if (lane <= NB2) {  // slot tables of panel p+1: (p+1) mod (NB+2-d)
                    const int v = cs[lane] + 1;
                    S.colslot[par ^ 1][lane] = (lane <= NB && v < NB2 - lane) ? v : 0;
                }
                PTL(1);
                bsync();
                PTL(2);
                // block row q = p+NB+1 is assembled for the slots that fall free: diagonal block into the spare slot of
                // the diagonal ring, the others and the right-hand-side block into the staging area -- all from the
                // row's record, which a bulk copy brought into shared memory several panels ago
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
        __syncthreads();

        // the partial sums of G of the six update warps
        if (warp < kPanelUpdW) {
            gacc += __shfl_xor_sync(kFull, gacc, 1);
            gacc += __shfl_xor_sync(kFull, gacc, 2);
            if (t == 0) S.red[warp * 8 + g] = gacc;
        }
        __syncthreads();
        // ---------------- observations: y from the last diagonal block, strains from the accumulated
        //                  products, h = von Mises at the two observed Gauss points (src/fem_postprocess.py:172-185)
        const double *stgl = lst + ((NQ - 1) & 1) * LPB;  // last panel: [c][k] = Minv[k][c]
        const double *rdl = S.rd[(NQ - 1) & 1];
        if (MODE > 0 && tid == 6 * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (warp == 0) {
            if (lane < 8) {
                double gs = 0.0;
                for (int w = 0; w < kPanelUpdW; ++w) gs += S.red[w * 8 + lane];
                S.G[lane] = gs;
            }
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
