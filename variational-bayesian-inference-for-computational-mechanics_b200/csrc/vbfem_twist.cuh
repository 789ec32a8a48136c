// vbfem_twist.cuh -- the production kernel for meshes whose band fits on chip
// (Cook 20x10: n = 440, half bandwidth 25): one CTA per Monte-Carlo sample, two
// resident CTAs per SM, the whole factor in shared memory.
//
//   all warps   : zero band, per-element Q4 Gauss-point kernels (material subroutine,
//                 B^T C B), colour-ordered scatter into the twisted band        (a)+(b)
//   warps 0, 1  : twisted LDL^T (vbfem_band.cuh), fused forward elimination of the load
//                 vector, back substitution, observation (y, von Mises h), adjoint
//                 right-hand side, adjoint forward/back substitution with the same
//                 factor                                                        (c)+(d)
//   all warps   : element-wise contraction -psi^T (dK/dp) u, chain rule to x    (e)
//
// Included by vbfem.cu after DevModel / Args / obs helpers are defined.
#pragma once
#include "vbfem_band.cuh"

namespace vbfem {

// x <- (L D L^T)^-1 x for the twisted factor.  fr = 0: top front, fr = 1: bottom front; both
// warps call this (named barrier 1, 64 threads, pairs them).
template <int B>
__device__ __forceinline__ void twist_back_solve(const DevModel &M, const double *band, double *vec, int fr,
                                                 int lane) {
    constexpr int P = B + 1;
    const int mid_end = M.pT + P;
    if (fr == 0) {
        front_scale<B>(band, vec, mid_end, lane);
        front_back_sweep<B>(band, vec, mid_end - 1, M.pT, lane);  // the shared middle first
        asm volatile("bar.sync 1, 64;" ::: "memory");
        front_back_sweep<B>(band, vec, M.pT - 1, 0, lane);
    } else {
        const double *bb = band + M.bandB_off;
        double *vb = vec + mid_end;
        front_scale<B>(bb, vb, M.nB, lane);
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (lane < P) vb[M.nB + lane] = vec[mid_end - 1 - lane];  // middle rows in mirrored order
        __syncwarp();
        front_apply_known<B>(bb, vb, M.nB, lane);
        front_back_sweep<B>(bb, vb, M.nB - 1, 0, lane);
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
}

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

template <int B, int NT>
__global__ void __launch_bounds__(NT, 2) fem_twist_kernel(const __grid_constant__ DevModel M,
                                                          const __grid_constant__ Args A) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_flag;
    constexpr int P = B + 1, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // n counts the padded system (dummy identity rows make pT and nB multiples of P); the P middle
    // rows follow the top front's pT columns; the bottom front's band is extended by P columns that
    // receive its Schur contribution to the middle.
    const int n = M.n, pT = M.pT, nB = M.nB;
    const int mid_end = pT + P, nvec = n + P;
    const int band_len = (n + P) * P;
    double *band = smem;
    double *bandB = band + M.bandB_off;
    double *vecU = smem + M.vec_off;   // rhs -> u, local vector order (top | bottom mirrored | scratch)
    double *vecP = vecU + nvec;        // adjoint rhs -> psi
    double *red = smem + M.red_off;
    double *obs_s = red + 2 * NW;
    // The two resident CTAs of an SM put their fronts on different scheduler partitions.
    const int fw = (NW >= 4) ? 2 * ((blockIdx.x / M.num_sms) & 1) : 0;
    const int fr = warp - fw;  // 0: top front, 1: bottom front, else helper

    for (long long s = blockIdx.x; s < A.N; s += gridDim.x) {
        // ---------------- sample parameters: theta -> (E, nu) -> (lambda, mu)
        // src/data_generation_2sam_more_loss.py:181-186
        double x0 = 0.0, x1 = 0.0, E, nu;
        if (A.mode & kElbo) {
            // main_custom_training.py:199-209: theta = e * sqrt(sig2) + mu, flattened [B*S]
            const long long j = A.j_begin + s;
            const int bb = (int)(j / A.S), ss = (int)(j % A.S);
            x0 = A.e[2 * ss] * sqrt(A.sig2[2 * bb]) + A.mu[2 * bb];
            x1 = A.e[2 * ss + 1] * sqrt(A.sig2[2 * bb + 1]) + A.mu[2 * bb + 1];
        } else if (A.x) {
            x0 = A.x[2 * s];
            x1 = A.x[2 * s + 1];
        }
        double *ws_s = A.ws ? A.ws + (size_t)s * A.ws_stride : nullptr;
        if (A.emat) {
            E = A.emat[2 * s];
            nu = A.emat[2 * s + 1];
        } else if (A.mode & kLoad) {
            E = ws_s[band_len + n];
            nu = ws_s[band_len + n + 1];
        } else {
            E = exp(M.theta_std[0] * x0 + M.theta_mean[0]);
            nu = 0.5 / (1.0 + exp(-M.theta_std[1] * x1 - M.theta_mean[1]));
        }
        const Lame mat = lame_from_E_nu(E, nu);
        if (tid == 0) s_flag = 0;

        if (!(A.mode & kLoad)) {
            // ---------------- zero the band, load the right-hand side, clear the adjoint vector
            {
                double2 *b2 = reinterpret_cast<double2 *>(band);
                const double2 z2 = make_double2(0.0, 0.0);
                for (int i = tid; i < band_len / 2; i += NT) b2[i] = z2;
                for (int i = tid; i < nvec; i += NT) {
                    vecU[i] = M.pf_loc[i];
                    vecP[i] = 0.0;
                }
            }
            __syncthreads();
            if (tid < M.ndummy) bandB[tid * P] = 1.0;  // padding rows: identity

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
                        // 36 offsets (16-bit, padded to 40) as five 128-bit loads
                        const uint4 *o4 = reinterpret_cast<const uint4 *>(M.eoff + 40 * e);
                        short off[40];
#pragma unroll
                        for (int q = 0; q < 5; ++q) *reinterpret_cast<uint4 *>(off + 8 * q) = o4[q];
#pragma unroll
                        for (int q = 0; q < 36; ++q)
                            if (off[q] >= 0) band[off[q]] += ke[q];
                    }
                    __syncthreads();
                }
            }
        } else {
            // ---------------- reload factor + solution kept by a previous forward launch
            for (int i = tid; i < band_len; i += NT) band[i] = ws_s[i];
            for (int i = tid; i < nvec; i += NT) {
                vecU[i] = (i < n) ? ws_s[band_len + i] : 0.0;
                vecP[i] = 0.0;
            }
            __syncthreads();
        }

        const bool adj = (A.mode & (kAdjoint | kLoad)) != 0;
        double f0 = 0.0, f1 = 0.0, gh0 = 0.0, gh1 = 0.0;
        if (fr == 0 || fr == 1) {
            double *bnd = fr ? bandB : band;
            if (!(A.mode & kLoad)) {
                // ---------------- (c) twisted LDL^T with fused forward elimination of the load vector
                FrontState<B> st;
                double *z = fr ? vecU + mid_end : vecU;
                front_init<B>(st, bnd, z, lane);
#pragma unroll 1
                for (int seg = 0; seg < 2; ++seg) {  // one copy of the column loop for all segments
                    if (seg == 1) {
                        if (fr == 1) {
                            front_flush<B>(st);
                            front_dump_middle<B>(st, bnd, nB, z);
                            __syncwarp();
                        }
                        asm volatile("bar.sync 1, 64;" ::: "memory");
                        if (fr == 1) break;
                        front_merge_middle<B>(st, bandB, nB, vecU + mid_end);
                    }
                    front_eliminate<B>(st, bnd, fr ? nB + P : mid_end, z, seg ? pT : 0,
                                       seg ? 1 : (fr ? nB : pT) / P);
                }
                if (st.bad < 0 && lane == 0) s_flag = 1;
                twist_back_solve<B>(M, band, vecU, fr, lane);
            }

            // ---------------- (d) observations: y = u(obs node), h = von Mises at (obs ele, obs gps)
            if (fr == 0) {
                if (lane < 2) {
                    double ue[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) ue[a] = (M.obs_lmv[a] >= 0) ? vecU[M.obs_lmv[a]] : 0.0;
                    double *o = obs_s + 12 * lane;
                    o[0] = obs_eval_pre(M, mat, ue, lane, adj ? o + 1 : nullptr, o + 9, o + 10);
                    if (A.h && !(A.mode & kLoad)) A.h[2 * s + lane] = o[0];
                }
                __syncwarp();
                f0 = (M.obs_lv[0] >= 0) ? vecU[M.obs_lv[0]] : 0.0;
                f1 = (M.obs_lv[1] >= 0) ? vecU[M.obs_lv[1]] : 0.0;
                // a zero or NaN pivot does not set a sign bit but poisons the solution
                if (lane == 0 && !(fabs(f0) < 1.0e300 && fabs(f1) < 1.0e300)) s_flag = 1;
                if (lane == 0 && !(A.mode & kLoad)) {
                    if (A.y) {
                        A.y[2 * s] = f0;
                        A.y[2 * s + 1] = f1;
                    }
                    if (A.f_out) {
                        A.f_out[2 * s] = f0;
                        A.f_out[2 * s + 1] = f1;
                    }
                }
            }
            if (adj) {
                // ---------------- adjoint right-hand side w = d(gy.y + gh.h)/du
                if (fr == 0) {
                    double gy0, gy1;
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
                    if (lane == 0) {
                        if (M.obs_lv[0] >= 0) vecP[M.obs_lv[0]] += gy0;
                        if (M.obs_lv[1] >= 0) vecP[M.obs_lv[1]] += gy1;
#pragma unroll
                        for (int a = 0; a < 8; ++a)
                            if (M.obs_lmv[a] >= 0)
                                vecP[M.obs_lmv[a]] += gh0 * obs_s[1 + a] + gh1 * obs_s[12 + 1 + a];
                        obs_s[24] = gh0;
                        obs_s[25] = gh1;
                    }
                    __syncwarp();
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");
                // ---------------- (e) K psi = w with the same factor: inward sweeps, middle, outward sweeps
                if (fr == 0)
                    front_fwd_sweep<B>(band, vecP, M.j0T, pT, mid_end, lane);
                else
                    front_fwd_sweep<B>(bandB, vecP + mid_end, M.j0B, nB, nB + P, lane);
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (fr == 0) {
                    if (lane < P) vecP[pT + lane] += vecP[mid_end + nB + P - 1 - lane];
                    __syncwarp();
                    front_fwd_sweep<B>(band, vecP, pT, mid_end, mid_end, lane);
                }
                twist_back_solve<B>(M, band, vecP, fr, lane);
            }
        }
        __syncthreads();

        if (A.mode & kKeep) {  // factor + solution + material parameters -> workspace slot of this sample
            for (int i = tid; i < band_len; i += NT) ws_s[i] = band[i];
            for (int i = tid; i < n; i += NT) ws_s[band_len + i] = vecU[i];
            if (tid == 0) {
                ws_s[band_len + n] = E;
                ws_s[band_len + n + 1] = nu;
            }
        }

        // ---------------- full fields for fem_test / fem_postprocess (src/fem_solver_tf.py:310-341)
        if (A.mode & kFields) {
            if (A.u_out) {
                for (int g = tid; g < M.ndof; g += NT) A.u_out[(size_t)s * M.ndof + g] = 0.0;
                __syncthreads();
                for (int r = tid; r < n; r += NT)
                    if (M.lv2dof[r] >= 0) A.u_out[(size_t)s * M.ndof + M.lv2dof[r]] = vecU[r];
            }
            if (A.fint_out) {
                for (int g = tid; g < M.ndof; g += NT) A.fint_out[(size_t)s * M.ndof + g] = 0.0;
                __syncthreads();
            }
            for (int base = 0; base < M.nele; base += NT) {
                const int k = base + tid;
                double p[8];
                int color = -1, e = 0;
                if (k < M.nele) {
                    e = M.eorder[k];
                    for (int c = 0; c < M.ncolors; ++c)
                        if (k >= M.color_start[c] && k < M.color_start[c + 1]) color = c;
                    double xl[4], yl[4], ue[8];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int nd = M.ien[4 * e + a];
                        xl[a] = M.coord[2 * nd];
                        yl[a] = M.coord[2 * nd + 1];
                    }
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const int r = M.ulm[8 * e + a];
                        ue[a] = (r >= 0) ? vecU[r] : 0.0;
                        p[a] = 0.0;
                    }
                    for (int gp = 0; gp < 4; ++gp) {
                        ShapeQ4 sh;
                        shapef_q4(xl, yl, gp, M.thk, sh);
                        double exx, eyy, gxy, sig[4];
                        Tangent C;
                        strain_q4(sh, ue, exx, eyy, gxy);
                        mat_isotropic_plane_strain(mat, exx, eyy, gxy, sig, C);
                        // p += dvol * Bm^T sig[0,1,3]   (src/mat_subroutine_tf.py:147-159)
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            p[2 * a] += sh.dvol * (sh.nx[a] * sig[0] + sh.ny[a] * sig[3]);
                            p[2 * a + 1] += sh.dvol * (sh.ny[a] * sig[1] + sh.nx[a] * sig[3]);
                        }
                        const size_t o = ((size_t)s * 6 * 4 + gp) * M.nele + e;  // [N][6][4][nele]
                        const size_t cs = (size_t)4 * M.nele;
                        if (A.sig_out) {
                            A.sig_out[o] = sig[0];
                            A.sig_out[o + cs] = sig[1];
                            A.sig_out[o + 2 * cs] = sig[2];
                            A.sig_out[o + 3 * cs] = sig[3];
                            A.sig_out[o + 4 * cs] = 0.0;
                            A.sig_out[o + 5 * cs] = 0.0;
                        }
                        if (A.eps_out) {
                            A.eps_out[o] = exx;
                            A.eps_out[o + cs] = eyy;
                            A.eps_out[o + 2 * cs] = 0.0;
                            A.eps_out[o + 3 * cs] = gxy;
                            A.eps_out[o + 4 * cs] = 0.0;
                            A.eps_out[o + 5 * cs] = 0.0;
                        }
                    }
                }
                if (A.fint_out) {
                    for (int c = 0; c < M.ncolors; ++c) {
                        if (color == c)
                            for (int a = 0; a < 8; ++a) A.fint_out[(size_t)s * M.ndof + M.lmg[8 * e + a]] += p[a];
                        __syncthreads();
                    }
                }
            }
        }

        // ---------------- element-wise contraction -psi^T (dK/dp) u + explicit dh/dp, chained to x
        if (adj) {
            double sl = 0.0, sm = 0.0;
            for (int k = tid; k < M.nele; k += NT) {
                const int e = M.eorder[k];
                double xl[4], yl[4], ue[8], pe[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int nd = M.ien[4 * e + a];
                    const double2 xy = *reinterpret_cast<const double2 *>(M.coord + 2 * nd);
                    xl[a] = xy.x;
                    yl[a] = xy.y;
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int r = M.ulm[8 * e + a];
                    ue[a] = (r >= 0) ? vecU[r] : 0.0;
                    pe[a] = (r >= 0) ? vecP[r] : 0.0;
                }
#pragma unroll 1
                for (int gp = 0; gp < 4; ++gp) {
                    ShapeQ4 sh;
                    shapef_q4(xl, yl, gp, M.thk, sh);
                    double uxx, uyy, uxy, pxx, pyy, pxy, cl, cm;
                    strain_q4(sh, ue, uxx, uyy, uxy);
                    strain_q4(sh, pe, pxx, pyy, pxy);
                    mat_tangent_param_contract(pxx, pyy, pxy, uxx, uyy, uxy, cl, cm);
                    sl = fma(sh.dvol, cl, sl);
                    sm = fma(sh.dvol, cm, sm);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sl += __shfl_down_sync(kFull, sl, o);
                sm += __shfl_down_sync(kFull, sm, o);
            }
            if (lane == 0) {
                red[2 * warp] = sl;
                red[2 * warp + 1] = sm;
            }
            __syncthreads();
            if (tid == 0) {
                double tl = 0.0, tm = 0.0;
                for (int w = 0; w < NW; ++w) {
                    tl += red[2 * w];
                    tm += red[2 * w + 1];
                }
                gh0 = obs_s[24];
                gh1 = obs_s[25];
                const double gl = -tl + gh0 * obs_s[9] + gh1 * obs_s[12 + 9];
                const double gm = -tm + gh0 * obs_s[10] + gh1 * obs_s[12 + 10];
                const double t = (1.0 + nu) * (1.0 - 2.0 * nu);
                const double dl_dE = mat.lam / E, dm_dE = mat.mu / E;
                const double dl_dnu = E * (1.0 + 2.0 * nu * nu) / (t * t);
                const double dm_dnu = -0.5 * E / ((1.0 + nu) * (1.0 + nu));
                const double gE = gl * dl_dE + gm * dm_dE;
                const double gnu = gl * dl_dnu + gm * dm_dnu;
                // dE/dx0 = std0 * E ; dnu/dx1 = std1 * nu (1 - 2 nu)
                A.gx[2 * s] = gE * M.theta_std[0] * E;
                A.gx[2 * s + 1] = gnu * M.theta_std[1] * nu * (1.0 - 2.0 * nu);
            }
        }
        if (tid == 0 && A.status && !(A.mode & kLoad)) A.status[s] = s_flag;
        __syncthreads();
    }
}

}  // namespace vbfem
