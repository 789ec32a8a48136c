// vbfem_front_kernel.cuh -- the production kernel for meshes whose band fits on chip
// (Cook 20x10: n = 440, half bandwidth 25): one CTA per Monte-Carlo sample, two resident CTAs
// per SM, the whole factor in shared memory.
//
//   all warps   : zero band, per-element Q4 Gauss-point kernels (material subroutine,
//                 B^T C B), colour-ordered scatter into the twisted band             (a)+(b)
//   warps f, f+1: twisted LDL^T (vbfem_front.cuh) with the load vector and the two unit
//                 vectors of the observed node eliminated on the fly; the observed
//                 displacement y is a dot product of eliminated vectors, the observed
//                 element lies in the middle block, so y and the von Mises measure h
//                 need NO back substitution                                          (c)+(d)
//                 forward+adjoint: the adjoint right-hand side is known at that point,
//                 so ONE back substitution carries u and psi together; Jacobian mode
//                 carries u and the four adjoint vectors of (y0, y1, h0, h1)
//   all warps   : element-wise contraction -psi^T (dK/dp) u, chain rule to x           (e)
//
// MODE 0: y, h                      (vbfem_forward)
// MODE 1: y, h, gx = J^T (gy, gh)   (vbfem_forward_backward, vbfem_elbo_step1)
// MODE 2: y, h, J = d(y, h)/dx      (vbfem_forward with keep_factor: vbfem_backward is J^T g)
//
// Included by vbfem.cu after DevModel / Args are defined.
#pragma once
#include "vbfem_front.cuh"

namespace vbfem {

// Observation at Gauss point slot q (0/1) of the observed element with host-precomputed
// shape-function derivatives (geometry is sample independent).
__device__ __forceinline__ double obs_eval_pre(const DevModel &M, const Lame &mat, const double (&ue)[8], int q,
                                               double *dhdu, double *dhdl, double *dhdm) {
    ShapeQ4 s;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        s.nx[a] = M.obs_nx[q][a];
        s.ny[a] = M.obs_ny[q][a];
    }
    s.dvol = 0.0;
    double exx, eyy, gxy;
    strain_q4(s, ue, exx, eyy, gxy);
    double sig[4];
    Tangent C;
    mat_isotropic_plane_strain(mat, exx, eyy, gxy, sig, C);
    double ds[4];
    const double h = von_mises_ref(sig, dhdu ? ds : nullptr);
    if (dhdu) {
        const double l2m = mat.lam + 2.0 * mat.mu;
        const double dexx = ds[0] * l2m + ds[1] * mat.lam + ds[2] * mat.lam;
        const double deyy = ds[0] * mat.lam + ds[1] * l2m + ds[2] * mat.lam;
        const double dgxy = ds[3] * mat.mu;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            dhdu[2 * a] = dexx * s.nx[a] + dgxy * s.ny[a];
            dhdu[2 * a + 1] = deyy * s.ny[a] + dgxy * s.nx[a];
        }
        *dhdl = (ds[0] + ds[1] + ds[2]) * (exx + eyy);
        *dhdm = 2.0 * ds[0] * exx + 2.0 * ds[1] * eyy + ds[3] * gxy;
    }
    return h;
}

template <int B, int NT, int MODE>
__global__ void __launch_bounds__(NT, 2) fem_front_kernel(const __grid_constant__ DevModel M,
                                                          const __grid_constant__ Args A) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_flag;
    constexpr int P = B + 1, NW = NT / 32;
    constexpr int NADJ = (MODE == 2) ? 4 : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = M.n, pT = M.pT, nB = M.nB, me = pT + P;
    // band: top front + middle columns [0, me), then the bottom front's nB columns (mirrored order)
    double *bandT = smem;
    double *bandB = smem + me * P;
    // five vectors of length n, local order [top | middle | bottom mirrored]; the internal numbering
    // ENDS at the observed node, so its two unit vectors live in the bottom front (three
    // right-hand sides there, one on the top front, which in turn finishes the middle block):
    //   X0 psi / adjoint of h0, X1 adjoint of h1, X2 load -> u, X3 / X4 unit vectors of the observed
    //   node -> their adjoints.  X0 and X1 are zero during the elimination: they back the bottom
    //   front's band (columns >= nB read as zero) and then carry the Schur hand-over.  The bottom
    //   front's right-hand-side reloads for its middle rows hit the (zero) top part of the next
    //   vector, resp. the zero pad behind X4.
    double *X = smem + n * P;
    double *X0 = X, *X1 = X + n, *X2 = X + 2 * n, *X3 = X + 3 * n, *X4 = X + 4 * n;
    double *zpad = X + 5 * n;  // 32 zeros
    double *obs_s = zpad + 32;  // 2 x 12 observation slots, then gy0, gy1, gh0, gh1
    double *red = obs_s + 32;   // 2 * NADJ * NW reduction slots
    // The two resident CTAs of an SM put their fronts on different scheduler partitions.  A warp's
    // partition is its hardware slot modulo 4 (measured, profiles/micro/warpid.cu: the second CTA of
    // an SM gets slots 5, 6, 7, 4); each CTA draws a ticket from its SM's counter (consecutive
    // tickets differ in parity) and takes partitions {0, 1} or {2, 3} for its two fronts.
    __shared__ int s_parts;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_flag = atomicAdd(M.sm_ticket + smid, 1);
        s_parts = 0;
    }
    __syncthreads();
    unsigned hw_slot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw_slot));
    if (lane == 0) atomicOr(&s_parts, 1 << (hw_slot & 3));
    __syncthreads();
    const int half = 2 * (s_flag & 1);
    // fall back to the warp index if the four warps do not sit on four different partitions
    const int fr = (NW == 4 && s_parts == 15) ? (int)(hw_slot & 3) - half : warp - half;  // 0 top, 1 bottom front
    __syncthreads();

    for (long long s = blockIdx.x; s < A.N; s += gridDim.x) {
        // ---------------- sample parameters: theta -> (E, nu) -> (lambda, mu)
        // src/data_generation_2sam_more_loss.py:181-186
        double x0 = 0.0, x1 = 0.0;
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
        const double E = exp(M.theta_std[0] * x0 + M.theta_mean[0]);
        const double nu = 0.5 / (1.0 + exp(-M.theta_std[1] * x1 - M.theta_mean[1]));
        const Lame mat = lame_from_E_nu(E, nu);
        if (tid == 0) s_flag = 0;
        VBFEM_TL(0);

        // ---------------- zero the band and the vectors, load the right-hand sides
        {
            double2 *b2 = reinterpret_cast<double2 *>(smem);
            const double2 z2 = make_double2(0.0, 0.0);
            const int nz = (n * P + 5 * n + 32) / 2;
            for (int i = tid; i < nz; i += NT) b2[i] = z2;
        }
        __syncthreads();
        VBFEM_TL(1);
        for (int i = tid; i < n; i += NT) X2[i] = M.pf_loc[i];
        if (tid < 2 && M.obs_lv[tid] >= 0) X[(3 + tid) * n + M.obs_lv[tid]] = 1.0;

        // ---------------- (a) element kernels + (b) colour-ordered scatter assembly
        for (int base = 0; base < M.nele; base += NT) {
            const int k = base + tid;
            double ke[36];
            int color = -1, e = 0;
            if (k < M.nele) {
                e = M.eorder[k];
#pragma unroll
                for (int c = 0; c < kMaxColors; ++c)
                    if (c < M.ncolors && k >= M.color_start[c] && k < M.color_start[c + 1]) color = c;
                double xl[4], yl[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int nd = M.ien[4 * e + a];
                    const double2 xy = *reinterpret_cast<const double2 *>(M.coord + 2 * nd);
                    xl[a] = xy.x;
                    yl[a] = xy.y;
                }
#pragma unroll
                for (int q = 0; q < 36; ++q) ke[q] = 0.0;
#pragma unroll 1
                for (int gp = 0; gp < 4; ++gp) {
                    ShapeQ4 sh;
                    shapef_q4(xl, yl, gp, M.thk, sh);
                    // zero predictor (src/fem_solver_tf.py:105-124): strain = 0, only the tangent matters
                    double sig[4];
                    Tangent C;
                    mat_isotropic_plane_strain(mat, 0.0, 0.0, 0.0, sig, C);
                    accumulate_kt(sh, C, ke);
                }
            }
            for (int c = 0; c < M.ncolors; ++c) {
                if (color == c) {
                    // 36 byte offsets into shared memory as nine 128-bit loads (entries of supported
                    // dofs point at a scratch slot); the 36 targets of one element are distinct and
                    // no other element of this colour shares them: load all, then store all
                    const uint4 *o4 = reinterpret_cast<const uint4 *>(M.eoff + 36 * e);
                    unsigned off[36];
#pragma unroll
                    for (int q = 0; q < 9; ++q) *reinterpret_cast<uint4 *>(off + 4 * q) = o4[q];
                    const unsigned sb = smem_addr(smem);
                    double cur[36];
#pragma unroll
                    for (int q = 0; q < 36; ++q)
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cur[q]) : "r"(sb + off[q]));
#pragma unroll
                    for (int q = 0; q < 36; ++q)
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(sb + off[q]), "d"(cur[q] + ke[q]) : "memory");
                }
                __syncthreads();
            }
        }

        VBFEM_TL(2);
        if (fr == 0 || fr == 1) {
            const unsigned vsb = 8u * (unsigned)n;  // right-hand sides X2, X3, X4 lie n doubles apart
            const unsigned S_sa = smem_addr(X0), rs_sa = S_sa + 8u * (P * P);
            double *yd = X0 + P * P + 3 * P + 2;  // bottom front's dot products (hand-over scratch)
            FrontState<B> st;
            if (fr == 0) {
                // ---------------- (c) top front: columns [0, pT) with the load vector, then the merged
                //                  middle block with all three right-hand sides
                const unsigned bT = smem_addr(bandT), zsT = smem_addr(X2);
                front_init<B, 1>(st, bT, zsT, vsb, lane);
                front_eliminate<B, 1>(st, bT, zsT, vsb, 0, pT);
                front_flush<B>(st);
                VBFEM_TL(3);
                asm volatile("bar.sync 1, 64;" ::: "memory");
                VBFEM_TL(4);
                front_merge_middle<B>(st, S_sa, rs_sa);
                const double ydB0 = yd[0], ydB1 = yd[1];
                front_eliminate<B, 3>(st, bT, zsT, vsb, pT, me);
                VBFEM_TL(5);
                // y = e_node^T K^-1 f = sum_j z_f z_e / d over the bottom front and the middle
                const double f0 = (M.obs_lv[0] >= 0) ? st.ydot[0] + ydB0 : 0.0;
                const double f1 = (M.obs_lv[1] >= 0) ? st.ydot[1] + ydB1 : 0.0;
                __syncwarp();
                // middle block of u: D^-1, then L_M^T
                const int bad = front_scale<B, 1>(bandT, X2, 0, 0, me, lane);
                if (bad || (lane == 0 && !(fabs(f0) < 1.0e300 && fabs(f1) < 1.0e300))) s_flag = 1;
                front_back_sweep<B, 1>(bandT, X2, 0, me - 1, pT, lane);
                // ---------------- (d) observations: y, h = von Mises at (obs ele, obs gps)
                if (lane < 2) {
                    double ue[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) ue[a] = (M.obs_lmv[a] >= 0) ? X2[M.obs_lmv[a]] : 0.0;
                    double *o = obs_s + 12 * lane;
                    o[0] = obs_eval_pre(M, mat, ue, lane, (MODE > 0) ? o + 1 : nullptr, o + 9, o + 10);
                    if (A.h) A.h[2 * s + lane] = o[0];
                }
                if (lane == 0) {
                    if (A.y) {
                        A.y[2 * s] = f0;
                        A.y[2 * s + 1] = f1;
                    }
                    if (A.f_out) {
                        A.f_out[2 * s] = f0;
                        A.f_out[2 * s + 1] = f1;
                    }
                }
                __syncwarp();
                VBFEM_TL(6);
                if (MODE == 1) {
                    // ---------------- adjoint right-hand side w = d(gy.y + gh.h)/du
                    double gy0, gy1, gh0 = 0.0, gh1 = 0.0;
                    if (A.mode & kElbo) {
                        // d(loss)/d f_j through term2 with the [B, B*S] broadcast (main_custom_training.py:205-214)
                        gy0 = A.gcoef * ((double)A.B * f0 - A.ysum[0]);
                        gy1 = A.gcoef * ((double)A.B * f1 - A.ysum[1]);
                    } else {
                        gy0 = A.gy[2 * s];
                        gy1 = A.gy[2 * s + 1];
                        gh0 = A.gh[2 * s];
                        gh1 = A.gh[2 * s + 1];
                    }
                    for (int r = lane; r < me; r += 32) X0[r] = 0.0;  // also clears the hand-over scratch
                    if (lane == 0) {
                        obs_s[24] = gy0;
                        obs_s[25] = gy1;
                        obs_s[26] = gh0;
                        obs_s[27] = gh1;
                    }
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int a = 0; a < 8; ++a)
                            if (M.obs_lmv[a] >= 0) X0[M.obs_lmv[a]] += gh0 * obs_s[1 + a] + gh1 * obs_s[12 + 1 + a];
                    }
                    __syncwarp();
                    // middle block of psi: L_M^-1 w_M joins the eliminated unit vectors, D^-1, L_M^T
                    front_fwd_sweep<B, 1>(bandT, X0, 0, pT, me, me, lane);
                    if (lane < P) {
                        const int r = pT + lane;
                        X0[r] = (X0[r] + gy0 * X3[r] + gy1 * X4[r]) * bandT[r * P];
                    }
                    __syncwarp();
                    front_back_sweep<B, 1>(bandT, X0, 0, me - 1, pT, lane);
                    VBFEM_TL(7);
                    asm volatile("bar.sync 1, 64;" ::: "memory");
                    VBFEM_TL(8);
                    front_back_sweep<B, 2>(bandT, X0, 2 * n, pT - 1, 0, lane);  // psi and u together
                    VBFEM_TL(9);
                } else if (MODE == 2) {
                    for (int r = lane; r < me; r += 32) X0[r] = X1[r] = 0.0;
                    __syncwarp();
                    if (lane < 2) {
#pragma unroll
                        for (int a = 0; a < 8; ++a)
                            if (M.obs_lmv[a] >= 0) X[lane * n + M.obs_lmv[a]] = obs_s[12 * lane + 1 + a];
                    }
                    __syncwarp();
                    front_fwd_sweep<B, 2>(bandT, X0, n, pT, me, me, lane);
                    front_scale<B, 2>(bandT, X0, n, pT, me, lane);
                    front_scale<B, 2>(bandT, X3, n, pT, me, lane);
                    front_back_sweep<B, 2>(bandT, X0, n, me - 1, pT, lane);
                    front_back_sweep<B, 2>(bandT, X3, n, me - 1, pT, lane);
                    asm volatile("bar.sync 1, 64;" ::: "memory");
                    front_back_sweep<B, 5>(bandT, X0, n, pT - 1, 0, lane);
                }
            } else {
                // ---------------- (c) bottom front: mirrored columns, the load vector and the observed
                //                  node's two unit vectors are eliminated on the fly
                const unsigned bBs = smem_addr(bandB), zsB = smem_addr(X2 + me);
                front_init<B, 3>(st, bBs, zsB, vsb, lane);
                front_eliminate<B, 3>(st, bBs, zsB, vsb, 0, nB);
                front_flush<B>(st);
                VBFEM_TL(3);
                front_dump_middle<B>(st, S_sa, rs_sa);
                if (lane == 0) {
                    yd[0] = st.ydot[0];
                    yd[1] = st.ydot[1];
                }
                __syncwarp();
                asm volatile("bar.sync 1, 64;" ::: "memory");
                VBFEM_TL(4);
                // D^-1 on the bottom parts of u and of the eliminated unit vectors (and the pivot check)
                if (front_scale<B, (MODE > 0 ? 3 : 0)>(bandB, X2 + me, n, 0, nB, lane)) s_flag = 1;
                if (MODE > 0) {
                    VBFEM_TL(7);
                    asm volatile("bar.sync 1, 64;" ::: "memory");
                    VBFEM_TL(8);
                    if (MODE == 1) {
                        // forward-eliminated adjoint right-hand side on the bottom front: gy . (unit vectors)
                        const double gy0 = obs_s[24], gy1 = obs_s[25];
                        for (int c = lane; c < nB; c += 32) X0[me + c] = gy0 * X3[me + c] + gy1 * X4[me + c];
                        __syncwarp();
                        front_apply_known<B, 2>(bandB, X0 + me, X0 + pT, 2 * n, nB, lane);
                        front_back_sweep<B, 2>(bandB, X0 + me, 2 * n, nB - 1, 0, lane);
                        VBFEM_TL(9);
                    } else {
                        for (int c = lane; c < nB; c += 32) X0[me + c] = X1[me + c] = 0.0;
                        __syncwarp();
                        front_apply_known<B, 5>(bandB, X0 + me, X0 + pT, n, nB, lane);
                        front_back_sweep<B, 5>(bandB, X0 + me, n, nB - 1, 0, lane);
                    }
                }
            }
        }
        __syncthreads();
        VBFEM_TL(10);

        // ---------------- (e) element-wise contraction -psi^T (dK/dp) u + explicit dh/dp, chained to x
        if (MODE > 0) {
            double sl[NADJ], sm[NADJ];
#pragma unroll
            for (int v = 0; v < NADJ; ++v) sl[v] = sm[v] = 0.0;
            for (int k = tid; k < M.nele; k += NT) {
                const int e = M.eorder[k];
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
                    lm[a] = M.ulm[8 * e + a];
                    ue[a] = (lm[a] >= 0) ? X2[lm[a]] : 0.0;
                }
#pragma unroll 1
                for (int gp = 0; gp < 4; ++gp) {
                    ShapeQ4 sh;
                    shapef_q4(xl, yl, gp, M.thk, sh);
                    double uxx, uyy, uxy;
                    strain_q4(sh, ue, uxx, uyy, uxy);
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double *pv = (MODE == 1) ? X0 : (v == 0 ? X0 : v == 1 ? X1 : v == 2 ? X3 : X4);
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
                    red[2 * (v * NW + warp)] = sl[v];
                    red[2 * (v * NW + warp) + 1] = sm[v];
                }
            }
            __syncthreads();
            if (tid == 0) {
                // d lambda, d mu / d(E, nu), then dE/dx0 = std0 * E ; dnu/dx1 = std1 * nu (1 - 2 nu)
                const double t = (1.0 + nu) * (1.0 - 2.0 * nu);
                const double dl_dE = mat.lam / E, dm_dE = mat.mu / E;
                const double dl_dnu = E * (1.0 + 2.0 * nu * nu) / (t * t);
                const double dm_dnu = -0.5 * E / ((1.0 + nu) * (1.0 + nu));
                const double dE_dx0 = M.theta_std[0] * E, dnu_dx1 = M.theta_std[1] * nu * (1.0 - 2.0 * nu);
                double tl[NADJ], tm[NADJ];
#pragma unroll
                for (int v = 0; v < NADJ; ++v) {
                    tl[v] = tm[v] = 0.0;
                    for (int w = 0; w < NW; ++w) {
                        tl[v] += red[2 * (v * NW + w)];
                        tm[v] += red[2 * (v * NW + w) + 1];
                    }
                }
                if (MODE == 1) {
                    const double gh0 = obs_s[26], gh1 = obs_s[27];
                    const double gl = -tl[0] + gh0 * obs_s[9] + gh1 * obs_s[12 + 9];
                    const double gm = -tm[0] + gh0 * obs_s[10] + gh1 * obs_s[12 + 10];
                    A.gx[2 * s] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                    A.gx[2 * s + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                } else {
                    // rows of J: h0, h1 (adjoints X0, X1 + explicit part), y0, y1 (adjoints X3, X4)
                    double *J = A.ws + (size_t)s * A.ws_stride;
#pragma unroll
                    for (int v = 0; v < NADJ; ++v) {
                        const double gl = -tl[v] + (v < 2 ? obs_s[12 * v + 9] : 0.0);
                        const double gm = -tm[v] + (v < 2 ? obs_s[12 * v + 10] : 0.0);
                        const int row = (v < 2) ? 2 + v : v - 2;  // storage order y0, y1, h0, h1
                        J[2 * row] = (gl * dl_dE + gm * dm_dE) * dE_dx0;
                        J[2 * row + 1] = (gl * dl_dnu + gm * dm_dnu) * dnu_dx1;
                    }
                }
            }
        }
        if (tid == 0 && A.status) A.status[s] = s_flag;
        __syncthreads();
        VBFEM_TL(11);
    }
}

// gx = J^T (gy, gh) with the Jacobians kept by a MODE 2 launch (vbfem_backward).
__global__ void jac_apply_kernel(long long N, const double *__restrict__ J, long long stride,
                                 const double *__restrict__ gy, const double *__restrict__ gh,
                                 double *__restrict__ gx) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double *j = J + s * stride;
    const double g0 = gy[2 * s], g1 = gy[2 * s + 1], g2 = gh[2 * s], g3 = gh[2 * s + 1];
    gx[2 * s] = g0 * j[0] + g1 * j[2] + g2 * j[4] + g3 * j[6];
    gx[2 * s + 1] = g0 * j[1] + g1 * j[3] + g2 * j[5] + g3 * j[7];
}

}  // namespace vbfem
