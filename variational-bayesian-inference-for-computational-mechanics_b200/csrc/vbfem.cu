// vbfem.cu -- libvbfem.so: batched Cook's-membrane FEM forward + adjoint for B200 (sm_100a).
//
// One CTA per Monte-Carlo sample (persistent, grid-stride over the batch).  Per sample:
//   (a) per-element Q4 Gauss-point kernel: shape functions, material subroutine
//       (stress + consistent tangent), element stiffness in registers
//       [src/mat_subroutine_tf.py:23-110, src/fem_preprocess.py:1223-1285]
//   (b) atomics-free scatter assembly, colour by colour, into a banded SPD matrix under an
//       internal bandwidth-minimising numbering
//       [replaces tf.scatter_nd into dense Kg, src/fem_solver_tf.py:336-341]
//   (c) banded LDL^T with the forward substitution fused into the column loop, back
//       substitution [replaces tf.linalg.solve, src/fem_solver_tf.py:137]
//   (d) fused displacement / von Mises observation [src/fem_postprocess.py:172-185]
//   (e) adjoint: reuse the factor for K psi = dJ/du, contract -psi^T (dK/dp) u element
//       by element, chain to x [what tape.gradient derives, main_custom_training.py:252-256]
// Four kernels: fem_warp_kernel (vbfem_warp.cuh; one warp per sample, the elimination
// window in registers, 8x8 blocks on FP64 tensor-core MMAs -- the production path for Cook 20x10),
// fem_panel_kernel (vbfem_panel.cuh; one CTA per sample, wide bands, factor streamed to HBM -- Cook 80x40),
// fem_front_kernel (vbfem_front_kernel.cuh; band on chip, two warp-synchronous column fronts) and fem_kernel
// below (any bandwidth, band in shared memory or HBM, full fields, plane stress, per-element materials).
// Paths are relative to nfeng2022/Variational-Bayesian-Inference-for-Computational-Mechanics.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>
#include <string>
#include <vector>

#include "../../include/vbfem.h"
#include "vbfem_math.cuh"

namespace vbfem {

// ------------------------------------------------------------------------------------------
// Device-side model description (passed by value as a __grid_constant__ kernel parameter).
// ------------------------------------------------------------------------------------------
constexpr int kMaxColors = 16;

struct DevModel {
    int n, b, ldb;          // order, half bandwidth, column stride (b + 2: diag, b sub-diagonals, rhs slot)
    int nele, nnodes, ndof;
    int ncolors, nitems;    // nitems = b(b+1)/2 trailing-update entries + b rhs entries per column step
    int band_in_smem;
    int vec_off, red_off;   // offsets (in doubles) of the two work vectors / reduction scratch in smem
    int ring_off, ring_w;   // band in HBM: shared-memory ring of ring_w band columns (factor window / sweep staging)
    int obs_ele, obs_gp[2]; // 0-based
    int obs_dof[2];         // band rows of the observed node's (x, y) dofs, -1 if supported
    int obs_lmb[8];         // band rows of the observed element's dofs
    int w_first;            // first band row where the adjoint right-hand side can be non-zero
    int color_start[kMaxColors + 1];
    double obs_x[4], obs_y[4];
    double thk, theta_mean[2], theta_std[2];
    int stype;            // 1 plane stress, 2 plane strain (section['stype'], model_property_cards.py:28)
    const double *coord;  // [nnodes][2]
    const int *ien;       // [nele][4] 0-based
    const short *lmb;     // [nele][8] band row of each element dof, -1 if supported
    const int *lmg;       // [nele][8] global dof (0-based)
    const int *eorder;    // elements sorted by colour
    const double *pf;     // [n] load vector in band order
    const int *band2dof;  // [n] band row -> global dof (0-based)
    // ---- on-chip front kernel (vbfem_front_kernel.cuh): top front [0, pT), P middle rows, bottom
    //      front mirrored; the internal numbering starts at the observed node
    int pT, nB, num_sms;
    int obs_lv[2];          // local-vector index of the observed node's dofs (bottom front), -1 if supported
    int obs_lmv[8];         // local-vector index of the observed element's dofs (middle block), -1 if supported
    double obs_nx[2][4], obs_ny[2][4];  // dN/dx, dN/dy at the two observed Gauss points
    const unsigned *eoff;   // [nele][36] shared-memory BYTE offset of each lower-triangle element entry (supported
                            //            dofs: a scratch slot)
    const short *ulm;       // [nele][8] local-vector index of each element dof, -1 if supported
    const double *pf_loc;   // [n] load vector in local-vector order
    int *sm_ticket;         // [num_sms] running CTA tickets per SM
};

enum : int {
    kKeep = 1,    // also store the 4x2 Jacobian d(y, h)/dx per sample (vbfem_forward with keep_factor)
    kAdjoint = 2, // run the adjoint in the same launch
    kFields = 8,  // write u / strain / stress / F_int
    kElbo = 16    // x from (mu, sig2, e); upstream gradient from the ELBO data term
};

struct Args {
    long long N;
    int mode;
    const double *x, *emat;
    int emat_per_ele;  // emat is [N][nele][2] (heterogeneous material, fields / forward only) instead of [N][2]
    double *y, *h;
    const double *gy, *gh;
    double *gx;
    double *ws;
    long long ws_stride;  // doubles per workspace slot
    int *status;
    double *u_out, *sig_out, *eps_out, *fint_out;
    // ELBO mode
    const double *mu, *sig2, *e, *ysum;
    int B, S;
    long long j_begin;
    double gcoef;  // 1 / (sig_e * B * (B*S))
    double *f_out;
    // generic kernel, Jacobian mode: constant cotangent (gy0, gy1, gh0, gh1) for every sample, gx written
    // to gx[s * gx_stride + gx_off + k] (one launch per row of the 4x2 Jacobian)
    int const_g;
    double gc[4];
    long long gx_stride;
    int gx_off;
    long long *timeline;  // profiling builds (-DVBFEM_TIMELINE): clock64 marks per CTA and warp
};

// ------------------------------------------------------------------------------------------
// Warp-level triangular sweeps (half bandwidth <= 31).  Lane l owns the row congruent to l
// modulo 32 inside a sliding 32-row window; one DFMA + one broadcast shuffle per row on the
// dependency chain, L entries are prefetched off the chain.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_back_sweep(const double *__restrict__ band, int n, int b, int ldb,
                                                const double *rhs, int rhs_stride, const double *scale,
                                                double *out, int lane) {
    // Solves L^T x = diag(scale) * rhs  (scale == nullptr: rhs already scaled).  out may alias rhs only
    // when rhs_stride == 1 (in-place on a work vector).
    int r = (n - 1) - (((n - 1) - lane) & 31);
    double acc = 0.0;
    if (r >= 0) acc = rhs[(size_t)r * rhs_stride] * (scale ? scale[(size_t)r * ldb] : 1.0);
#pragma unroll 4
    for (int j = n - 1; j >= 0; --j) {
        const int i = (j - lane) & 31;
        double lv = 0.0;
        if (i >= 1 && i <= b && r >= 0) lv = band[r * ldb + i];
        const double xj = __shfl_sync(0xffffffffu, acc, j & 31);
        acc = fma(-lv, xj, acc);
        if (i == 0) {
            out[j] = xj;
            r -= 32;
            acc = 0.0;
            if (r >= 0) acc = rhs[(size_t)r * rhs_stride] * (scale ? scale[(size_t)r * ldb] : 1.0);
        }
    }
}

__device__ __forceinline__ void warp_fwd_sweep(const double *__restrict__ band, int n, int b, int ldb,
                                               double *__restrict__ vec, int j0, int lane) {
    // In-place L z = w on vec, starting at row j0 (rows before j0 hold zeros).
    int r = j0 + ((lane - j0) & 31);
    double acc = (r < n) ? vec[r] : 0.0;
#pragma unroll 4
    for (int j = j0; j < n; ++j) {
        const int i = (lane - j) & 31;
        double lv = 0.0;
        if (i >= 1 && i <= b && r < n) lv = band[j * ldb + i];
        const double zj = __shfl_sync(0xffffffffu, acc, j & 31);
        acc = fma(-lv, zj, acc);
        if (i == 0) {
            vec[j] = zj;
            r += 32;
            acc = (r < n) ? vec[r] : 0.0;
        }
    }
}

// Observation at one Gauss point of the observed element: von Mises measure and, optionally,
// its derivatives w.r.t. the element displacements and the Lame parameters at fixed u.
__device__ __forceinline__ double obs_eval(const DevModel &M, const Lame &mat, const double (&ue)[8], int gp,
                                           double *dhdu /*[8]*/, double *dhdl, double *dhdm) {
    ShapeQ4 s;
    shapef_q4(M.obs_x, M.obs_y, gp, M.thk, s);
    double exx, eyy, gxy;
    strain_q4(s, ue, exx, eyy, gxy);
    double sig[4];
    Tangent C;
    mat_isotropic_plane_strain(mat, exx, eyy, gxy, sig, C);
    double ds[4];
    const double h = von_mises_ref(sig, dhdu ? ds : nullptr);
    if (dhdu) {
        const double l2m = mat.lam + 2.0 * mat.mu;
        const double dexx = ds[0] * l2m + ds[1] * mat.lam + ds[2] * mat.szz;
        const double deyy = ds[0] * mat.lam + ds[1] * l2m + ds[2] * mat.szz;
        const double dgxy = ds[3] * mat.mu;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            dhdu[2 * a] = dexx * s.nx[a] + dgxy * s.ny[a];
            dhdu[2 * a + 1] = deyy * s.ny[a] + dgxy * s.nx[a];
        }
        const double tr = exx + eyy;
        *dhdl = (ds[0] + ds[1] + (mat.szz != 0.0 ? ds[2] : 0.0)) * tr;  // sigma_zz carries lambda only in plane strain
        *dhdm = 2.0 * ds[0] * exx + 2.0 * ds[1] * eyy + ds[3] * gxy;
    }
    return h;
}

#ifdef VBFEM_TIMELINE
#define VBFEM_TL(i)                                                                              \
    do {                                                                                        \
        if (A.timeline && (threadIdx.x & 31) == 0)                                              \
            A.timeline[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 16 + (i)] = clock64();          \
    } while (0)
#else
#define VBFEM_TL(i) ((void)0)
#endif
}  // namespace vbfem
#include "vbfem_front_kernel.cuh"
#include "vbfem_panel.cuh"
#include "vbfem_panel2.cuh"
#include "vbfem_warp.cuh"
#include "vbfem_warp2.cuh"
#include "vbfem_peer.cuh"
namespace vbfem {

// ------------------------------------------------------------------------------------------
// Band-in-HBM helpers of the generic kernel (wide bands, e.g. Cook 80x40: n = 6560, b = 85).
// ------------------------------------------------------------------------------------------
constexpr int kRingDepth = 8;  // band columns fetched ahead of the factorisation window (cp.async groups in flight)

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Triangular sweeps with the factor in HBM: the band columns pass through the shared-memory ring in
// two halves of `ch` columns; warps 1.. stage the next half while warp 0 sweeps the current one.
//   DIR = -1: back substitution L^T x = y, column oriented: x_j = y_j - sum_i L[j+i][j] x_{j+i}
//   DIR = +1: forward substitution L z = w from column j0: z[j+i] -= L[j+i][j] z_j
template <int NT, int DIR>
__device__ __forceinline__ void staged_sweep(const double *__restrict__ band, double *__restrict__ ring, int ch,
                                             int n, int b, int ldb, double *__restrict__ x, int j0, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int first = (DIR > 0) ? j0 : n - 1;
    const int ncol = (DIR > 0) ? n - j0 : n;
    const int nchunk = (ncol + ch - 1) / ch;
    auto chunk_lo = [&](int q) { return (DIR > 0) ? first + q * ch : max(first - (q + 1) * ch + 1, 0); };
    auto chunk_cnt = [&](int q) { return (DIR > 0) ? min(ch, n - (first + q * ch)) : (first - q * ch) - chunk_lo(q) + 1; };
    auto stage = [&](int q, int t0, int nthr) {
        const int lo = chunk_lo(q), cnt = chunk_cnt(q);
        double *dst = ring + (q & 1) * ch * ldb;
        const double *src = band + (size_t)lo * ldb;
        for (int i = t0; i < cnt * ldb; i += nthr) dst[i] = src[i];
    };
    if (nchunk <= 0) return;
    stage(0, tid, NT);
    __syncthreads();
    for (int q = 0; q < nchunk; ++q) {
        if (warp > 0) {
            if (q + 1 < nchunk) stage(q + 1, tid - 32, NT - 32);
        } else {
            const int lo = chunk_lo(q), cnt = chunk_cnt(q);
            const double *cols = ring + (q & 1) * ch * ldb;
            if (DIR < 0) {
                int j = lo + cnt - 1;
                // four columns per trip: four independent partial dot products over the already final
                // entries (their shuffle reductions overlap), then the 4x4 coupling inside the block
                for (; j - 3 >= lo; j -= 4) {
                    double acc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double *c = cols + (j - k - lo) * ldb;
                        acc[k] = 0.0;
                        for (int i = k + 1 + lane; i <= b; i += 32)
                            if (j - k + i < n) acc[k] = fma(c[i], x[j - k + i], acc[k]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                    const double *c1 = cols + (j - 1 - lo) * ldb, *c2 = cols + (j - 2 - lo) * ldb,
                                 *c3 = cols + (j - 3 - lo) * ldb;
                    const double x0 = x[j] - acc[0];
                    const double x1 = fma(-c1[1], x0, x[j - 1] - acc[1]);
                    const double x2 = fma(-c2[1], x1, fma(-c2[2], x0, x[j - 2] - acc[2]));
                    const double x3 = fma(-c3[1], x2, fma(-c3[2], x1, fma(-c3[3], x0, x[j - 3] - acc[3])));
                    __syncwarp();
                    if (lane == 0) {
                        x[j] = x0;
                        x[j - 1] = x1;
                        x[j - 2] = x2;
                        x[j - 3] = x3;
                    }
                    __syncwarp();
                }
                for (; j >= lo; --j) {
                    const double *c = cols + (j - lo) * ldb;
                    double acc = 0.0;
                    for (int i = 1 + lane; i <= b; i += 32)
                        if (j + i < n) acc = fma(c[i], x[j + i], acc);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) x[j] -= acc;
                    __syncwarp();
                }
            } else {
                int j = lo;
                // four columns per trip: resolve the block's four unknowns (every lane redundantly),
                // then one pass over the rows below applies all four columns
                for (; j + 3 < lo + cnt; j += 4) {
                    const double *c0 = cols + (j - lo) * ldb, *c1 = c0 + ldb, *c2 = c1 + ldb, *c3 = c2 + ldb;
                    const double z0 = x[j];
                    const double z1 = fma(-c0[1], z0, x[j + 1]);
                    const double z2 = fma(-c1[1], z1, fma(-c0[2], z0, x[j + 2]));
                    const double z3 = fma(-c2[1], z2, fma(-c1[2], z1, fma(-c0[3], z0, x[j + 3])));
                    __syncwarp();
                    if (lane == 0) {
                        x[j + 1] = z1;
                        x[j + 2] = z2;
                        x[j + 3] = z3;
                    }
                    // rows r = j+4 .. j+3+b: entry of column j+k at offset r - (j+k), valid while <= b
                    for (int r = j + 4 + lane; r <= j + 3 + b && r < n; r += 32) {
                        const int o = r - j;
                        double a = x[r];
                        if (o <= b) a = fma(-c0[o], z0, a);
                        if (o - 1 <= b) a = fma(-c1[o - 1], z1, a);
                        if (o - 2 <= b) a = fma(-c2[o - 2], z2, a);
                        a = fma(-c3[o - 3], z3, a);
                        x[r] = a;
                    }
                    __syncwarp();
                }
                for (; j < lo + cnt; ++j) {
                    const double *c = cols + (j - lo) * ldb;
                    const double zj = x[j];
                    for (int i = 1 + lane; i <= b; i += 32)
                        if (j + i < n) x[j + i] = fma(-c[i], zj, x[j + i]);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Generic per-sample kernel (any bandwidth; band in shared memory if it fits, else in HBM).
// Used when the on-chip two-front kernel does not apply (e.g. the 80x40 mesh) and for full fields.
// ------------------------------------------------------------------------------------------
template <int NT, int EPT, int MINB, bool HBM>
__global__ void __launch_bounds__(NT, MINB) fem_kernel(const __grid_constant__ DevModel M,
                                                 const __grid_constant__ Args A) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int n = M.n, b = M.b, ldb = M.ldb;
    const int band_len = n * ldb;
    double *vec_u = smem + M.vec_off;  // solution u (band order)
    double *vec_p = vec_u + n;         // adjoint work vector
    double *red = smem + M.red_off;    // 2*NW reduction slots + 32 observation slots
    double *obs_s = red + 2 * NW;

    // --- static per-thread work items of one column step: A[tgt] -= col[s1] * (col[s2] / d)
    int it_tgt[EPT], it_s1[EPT], it_s2[EPT], it_row[EPT];
    {
        const int ntri = b * (b + 1) / 2;
#pragma unroll
        for (int k = 0; k < (HBM ? 0 : EPT); ++k) {
            int idx = tid + k * NT;
            it_tgt[k] = -1;
            it_s1[k] = it_s2[k] = it_row[k] = 0;
            if (idx < ntri) {
                int m = 1;
                while (idx >= b - m + 1) {
                    idx -= b - m + 1;
                    ++m;
                }
                it_tgt[k] = m * ldb + idx;  // A[j+m+t][j+m]
                it_s1[k] = m + idx;
                it_s2[k] = m;
                it_row[k] = m + idx;
            } else if (idx < ntri + b) {
                const int i = idx - ntri + 1;  // rhs slot of column j+i
                it_tgt[k] = i * ldb + (b + 1);
                it_s1[k] = i;
                it_s2[k] = b + 1;
                it_row[k] = i;
            }
        }
    }

    // band in HBM: the factorisation window lives in a shared-memory ring
    double *ring = smem + M.ring_off;
    const int ring_w = M.ring_w, ring_ch = M.ring_w / 2;

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
        double *ws_s = A.ws ? A.ws + (size_t)blockIdx.x * A.ws_stride : nullptr;  // per-CTA scratch (band in HBM)
        if (A.emat && !A.emat_per_ele) {
            E = A.emat[2 * s];
            nu = A.emat[2 * s + 1];
        } else {
            E = exp(M.theta_std[0] * x0 + M.theta_mean[0]);
            nu = 0.5 / (1.0 + exp(-M.theta_std[1] * x1 - M.theta_mean[1]));
        }
        auto make_mat = [&](double Ee, double ve) { return M.stype == 1 ? lame_plane_stress(Ee, ve) : lame_from_E_nu(Ee, ve); };
        // heterogeneous material: (E, nu) of element e of this sample; the observation uses the observed element's
        auto ele_mat = [&](int e) {
            const double *em = A.emat + ((size_t)s * M.nele + e) * 2;
            return make_mat(em[0], em[1]);
        };
        if (A.emat && A.emat_per_ele) {
            E = A.emat[((size_t)s * M.nele + M.obs_ele) * 2];
            nu = A.emat[((size_t)s * M.nele + M.obs_ele) * 2 + 1];
        }
        const Lame mat = make_mat(E, nu);
        double *band = M.band_in_smem ? smem : ws_s;
        if (tid == 0) s_flag = 0;

        {
            // ---------------- zero the band, load the right-hand side
            for (int i = tid; i < band_len; i += NT) band[i] = 0.0;
            __syncthreads();
            for (int r = tid; r < n; r += NT) band[r * ldb + (b + 1)] = M.pf[r];

            // ---------------- (a) element kernels + (b) coloured scatter assembly
            for (int base = 0; base < M.nele; base += NT) {
                const int k = base + tid;
                double ke[36];
                int lm[8];
                int color = -1;
                if (k < M.nele) {
                    const int e = M.eorder[k];
#pragma unroll
                    for (int c = 0; c < M.ncolors && c < kMaxColors; ++c)
                        if (k >= M.color_start[c] && k < M.color_start[c + 1]) color = c;
                    double xl[4], yl[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int nd = M.ien[4 * e + a];
                        xl[a] = M.coord[2 * nd];
                        yl[a] = M.coord[2 * nd + 1];
                    }
#pragma unroll
                    for (int a = 0; a < 8; ++a) lm[a] = M.lmb[8 * e + a];
#pragma unroll
                    for (int q = 0; q < 36; ++q) ke[q] = 0.0;
#pragma unroll 1
                    for (int gp = 0; gp < 4; ++gp) {
                        ShapeQ4 sh;
                        shapef_q4(xl, yl, gp, M.thk, sh);
                        // zero predictor (src/fem_solver_tf.py:105-124): strain = 0, only the tangent matters
                        double sig[4];
                        Tangent C;
                        mat_isotropic_plane_strain(A.emat_per_ele ? ele_mat(e) : mat, 0.0, 0.0, 0.0, sig, C);
                        accumulate_kt(sh, C, ke);
                    }
                }
                for (int c = 0; c < M.ncolors; ++c) {
                    if (color == c) {
                        if (HBM) {
                            // band in HBM: the 36 targets of an element are distinct and no other element of
                            // this colour shares them -- issue all loads, then all stores (a `+=` loop would
                            // serialise 36 HBM round trips on the assumed aliasing)
                            double cur[36];
#pragma unroll
                            for (int a = 0; a < 8; ++a)
#pragma unroll
                                for (int q = 0; q <= a; ++q) {
                                    const int pa = lm[a], pq = lm[q];
                                    const int lo = min(pa, pq), hi = max(pa, pq);
                                    cur[tri(a, q)] = (lo >= 0) ? band[(size_t)lo * ldb + (hi - lo)] : 0.0;
                                }
#pragma unroll
                            for (int a = 0; a < 8; ++a)
#pragma unroll
                                for (int q = 0; q <= a; ++q) {
                                    const int pa = lm[a], pq = lm[q];
                                    const int lo = min(pa, pq), hi = max(pa, pq);
                                    if (lo >= 0) band[(size_t)lo * ldb + (hi - lo)] = cur[tri(a, q)] + ke[tri(a, q)];
                                }
                        } else {
#pragma unroll
                            for (int a = 0; a < 8; ++a) {
#pragma unroll
                                for (int q = 0; q <= a; ++q) {
                                    const int pa = lm[a], pq = lm[q];
                                    if (pa >= 0 && pq >= 0) {
                                        const int lo = min(pa, pq), hi = max(pa, pq);
                                        band[lo * ldb + (hi - lo)] += ke[tri(a, q)];
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
            }

            // ---------------- (c) banded LDL^T, forward substitution fused (rhs slot b+1)
            if (!HBM) {
                for (int j = 0; j < n; ++j) {
                    double *colj = band + j * ldb;
                    const double d = colj[0];
                    if (tid == 0 && !(d > 0.0 && d < 1.0e300)) s_flag = 1;
                    const double rd = fast_rcp(d);
#pragma unroll
                    for (int k = 0; k < EPT; ++k) {
                        if (it_tgt[k] >= 0 && j + it_row[k] < n) {
                            const double v = colj[it_s1[k]];
                            const double w = colj[it_s2[k]] * rd;
                            colj[it_tgt[k]] = fma(-v, w, colj[it_tgt[k]]);
                        }
                    }
                    __syncthreads();
                }
                // L = V D^-1 (unit lower), diagonal slot <- 1/d, rhs slot <- D^-1 z
                for (int c = tid; c < n; c += NT) band[c * ldb] = fast_rcp(band[c * ldb]);
                __syncthreads();
                for (int c = warp; c < n; c += NW) {
                    const double rdc = band[c * ldb];
                    for (int i = 1 + lane; i <= b + 1; i += 32) band[c * ldb + i] *= rdc;
                }
                __syncthreads();
            } else {
                // Band in HBM: the b+1 columns under the trailing update live in a shared-memory ring
                // (column c <-> slot c mod ring_w).  While the block updates the window of column j, the
                // first ldb threads write the finished column j-1 back (already scaled: L = V D^-1, 1/d,
                // D^-1 z) and fetch column j-1+ring_w into its slot with cp.async, kRingDepth columns
                // ahead of its first use.
                __syncthreads();  // assembly complete in HBM
                for (int i = tid; i < min(ring_w, n) * ldb; i += NT) ring[i] = band[i];
                __syncthreads();
                int js = 0;
                double rd_prev = 0.0;
                for (int j = 0; j < n; ++j) {
                    const double *colj = ring + js * ldb;
                    const double d = colj[0];
                    if (tid == 0 && !(d > 0.0 && d < 1.0e300)) s_flag = 1;
                    const double rd = fast_rcp(d);
                    const double zjs = colj[b + 1];
                    // trailing update, one warp per target column j+m, lanes over its entries (no index
                    // tables here: the item arrays of the shared-memory path cost registers)
                    for (int m = 1 + warp; m <= b && j + m < n; m += NW) {
                        int slot = js + m;
                        slot -= (slot >= ring_w) ? ring_w : 0;
                        double *tc = ring + slot * ldb + lane;
                        const double *cs = colj + m + lane;
                        const double wm = colj[m] * rd;    // L[j+m][j]
                        const int cnt = b - m + 1 - lane;  // this lane's entries t = lane, lane+32, ... below cnt
                        if (b <= 127) {
                            // all loads first, then the FMAs, then the stores: source and target live in the
                            // same ring, a load-FMA-store loop would serialise on the assumed aliasing
                            double a[4], c[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                a[q] = (32 * q < cnt) ? tc[32 * q] : 0.0;
                                c[q] = (32 * q < cnt) ? cs[32 * q] : 0.0;
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (32 * q < cnt) tc[32 * q] = fma(-c[q], wm, a[q]);
                        } else {
                            for (int t = 0; t < cnt; t += 32) tc[t] = fma(-cs[t], wm, tc[t]);
                        }
                    }
                    // right-hand side slots of the b columns below: rhs[j+m] -= L[j+m][j] z_j, lanes over m
                    // (one dependent read-modify-write per target column inside the loop above would stall it)
                    if (warp == NW - 1) {
                        for (int m = 1 + lane; m <= b && j + m < n; m += 32) {
                            int slot = js + m;
                            slot -= (slot >= ring_w) ? ring_w : 0;
                            double *tr = ring + slot * ldb + (b + 1);
                            *tr = fma(-zjs, colj[m] * rd, *tr);
                        }
                    }
                    if (tid < ldb && j >= 1) {
                        int sp = js - 1;
                        sp += (sp < 0) ? ring_w : 0;
                        const double v = ring[sp * ldb + tid];
                        band[(size_t)(j - 1) * ldb + tid] = (tid == 0) ? rd_prev : v * rd_prev;
                        const int jn = j - 1 + ring_w;
                        if (jn < n) cp_async8(ring + sp * ldb + tid, band + (size_t)jn * ldb + tid);
                    }
                    cp_async_commit();
                    cp_async_wait<kRingDepth - 1>();
                    __syncthreads();
                    rd_prev = rd;
                    js = (js + 1 == ring_w) ? 0 : js + 1;
                }
                if (tid < ldb) {
                    int sp = js - 1;
                    sp += (sp < 0) ? ring_w : 0;
                    const double v = ring[sp * ldb + tid];
                    band[(size_t)(n - 1) * ldb + tid] = (tid == 0) ? rd_prev : v * rd_prev;
                }
                cp_async_wait<0>();
                __syncthreads();
            }

            // ---------------- back substitution  L^T u = D^-1 z
            if (b <= 31) {
                if (warp == 0) warp_back_sweep(band, n, b, ldb, band + (b + 1), ldb, nullptr, vec_u, lane);
            } else if (!HBM) {
                for (int r = tid; r < n; r += NT) vec_u[r] = band[r * ldb + (b + 1)];
                __syncthreads();
                for (int j = n - 1; j >= 0; --j) {
                    const double xj = vec_u[j];
                    for (int i = 1 + tid; i <= b; i += NT)
                        if (j - i >= 0) vec_u[j - i] = fma(-band[(j - i) * ldb + i], xj, vec_u[j - i]);
                    __syncthreads();
                }
            } else {
                for (int r = tid; r < n; r += NT) vec_u[r] = band[(size_t)r * ldb + (b + 1)];
                __syncthreads();
                staged_sweep<NT, -1>(band, ring, ring_ch, n, b, ldb, vec_u, 0, tid);
            }
            __syncthreads();
        }

        // ---------------- (d) observations: y = u(obs node), h = von Mises at (obs ele, obs gps)
        const bool adj = (A.mode & kAdjoint) != 0;
        if (lane == 0 && warp < 2) {
            double ue[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) ue[a] = (M.obs_lmb[a] >= 0) ? vec_u[M.obs_lmb[a]] : 0.0;
            double *o = obs_s + 12 * warp;
            o[0] = obs_eval(M, mat, ue, M.obs_gp[warp], adj ? o + 1 : nullptr, o + 9, o + 10);
            if (A.h) A.h[2 * s + warp] = o[0];
        }
        const double f0 = (M.obs_dof[0] >= 0) ? vec_u[M.obs_dof[0]] : 0.0;
        const double f1 = (M.obs_dof[1] >= 0) ? vec_u[M.obs_dof[1]] : 0.0;
        if (tid == 0) {
            if (A.y) {
                A.y[2 * s] = f0;
                A.y[2 * s + 1] = f1;
            }
            if (A.f_out) {
                A.f_out[2 * s] = f0;
                A.f_out[2 * s + 1] = f1;
            }
        }

        // ---------------- full fields for fem_test / fem_postprocess (src/fem_solver_tf.py:310-341)
        if (A.mode & kFields) {
            if (A.u_out) {
                for (int g = tid; g < M.ndof; g += NT) A.u_out[(size_t)s * M.ndof + g] = 0.0;
                __syncthreads();
                for (int r = tid; r < n; r += NT) A.u_out[(size_t)s * M.ndof + M.band2dof[r]] = vec_u[r];
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
                        const int r = M.lmb[8 * e + a];
                        ue[a] = (r >= 0) ? vec_u[r] : 0.0;
                        p[a] = 0.0;
                    }
                    for (int gp = 0; gp < 4; ++gp) {
                        ShapeQ4 sh;
                        shapef_q4(xl, yl, gp, M.thk, sh);
                        double exx, eyy, gxy, sig[4];
                        Tangent C;
                        strain_q4(sh, ue, exx, eyy, gxy);
                        mat_isotropic_plane_strain(A.emat_per_ele ? ele_mat(e) : mat, exx, eyy, gxy, sig, C);
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
                            // plane stress: eps33 = -v / (1 - v) (eps_xx + eps_yy)  (src/mat_subroutine.py:289, :51-52)
                            double e33 = 0.0;
                            if (M.stype == 1) {
                                const double ve = A.emat_per_ele ? A.emat[((size_t)s * M.nele + e) * 2 + 1] : nu;
                                e33 = -ve / (1.0 - ve) * (exx + eyy);
                            }
                            A.eps_out[o + 2 * cs] = e33;
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

        // ---------------- (e) adjoint: K psi = dJ/du with the same factor, then the element-wise
        //                  contraction -psi^T (dK/dp) u and the explicit dh/dp term, chained to x
        if (adj) {
            __syncthreads();  // obs_s complete
            double gy0, gy1, gh0 = 0.0, gh1 = 0.0;
            if (A.mode & kElbo) {
                // d(loss)/d f_j through term2 with the [B, B*S] broadcast (main_custom_training.py:205-214)
                gy0 = A.gcoef * ((double)A.B * f0 - A.ysum[0]);
                gy1 = A.gcoef * ((double)A.B * f1 - A.ysum[1]);
            } else if (A.const_g) {
                gy0 = A.gc[0];
                gy1 = A.gc[1];
                gh0 = A.gc[2];
                gh1 = A.gc[3];
            } else {
                gy0 = A.gy[2 * s];
                gy1 = A.gy[2 * s + 1];
                gh0 = A.gh[2 * s];
                gh1 = A.gh[2 * s + 1];
            }
            for (int r = tid; r < n; r += NT) vec_p[r] = 0.0;
            __syncthreads();
            if (tid == 0) {
                if (M.obs_dof[0] >= 0) vec_p[M.obs_dof[0]] += gy0;
                if (M.obs_dof[1] >= 0) vec_p[M.obs_dof[1]] += gy1;
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (M.obs_lmb[a] >= 0) vec_p[M.obs_lmb[a]] += gh0 * obs_s[1 + a] + gh1 * obs_s[12 + 1 + a];
            }
            __syncthreads();
            if (b <= 31) {
                if (warp == 0) {
                    warp_fwd_sweep(band, n, b, ldb, vec_p, M.w_first, lane);
                    __syncwarp();
                    warp_back_sweep(band, n, b, ldb, vec_p, 1, band, vec_p, lane);
                }
            } else if (HBM) {
                staged_sweep<NT, +1>(band, ring, ring_ch, n, b, ldb, vec_p, M.w_first, tid);
                for (int r = tid; r < n; r += NT) vec_p[r] *= band[(size_t)r * ldb];
                __syncthreads();
                staged_sweep<NT, -1>(band, ring, ring_ch, n, b, ldb, vec_p, 0, tid);
            } else {
                for (int j = M.w_first; j < n; ++j) {
                    const double zj = vec_p[j];
                    for (int i = 1 + tid; i <= b; i += NT)
                        if (j + i < n) vec_p[j + i] = fma(-band[j * ldb + i], zj, vec_p[j + i]);
                    __syncthreads();
                }
                for (int r = tid; r < n; r += NT) vec_p[r] *= band[r * ldb];
                __syncthreads();
                for (int j = n - 1; j >= 0; --j) {
                    const double xj = vec_p[j];
                    for (int i = 1 + tid; i <= b; i += NT)
                        if (j - i >= 0) vec_p[j - i] = fma(-band[(j - i) * ldb + i], xj, vec_p[j - i]);
                    __syncthreads();
                }
            }
            __syncthreads();
            double sl = 0.0, sm = 0.0;
            for (int k = tid; k < M.nele; k += NT) {
                const int e = M.eorder[k];
                double xl[4], yl[4], ue[8], pe[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int nd = M.ien[4 * e + a];
                    xl[a] = M.coord[2 * nd];
                    yl[a] = M.coord[2 * nd + 1];
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int r = M.lmb[8 * e + a];
                    ue[a] = (r >= 0) ? vec_u[r] : 0.0;
                    pe[a] = (r >= 0) ? vec_p[r] : 0.0;
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
                sl += __shfl_down_sync(0xffffffffu, sl, o);
                sm += __shfl_down_sync(0xffffffffu, sm, o);
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
                const double gl = -tl + gh0 * obs_s[9] + gh1 * obs_s[12 + 9];
                const double gm = -tm + gh0 * obs_s[10] + gh1 * obs_s[12 + 10];
                const double t = (1.0 + nu) * (1.0 - 2.0 * nu);
                const double dl_dE = mat.lam / E, dm_dE = mat.mu / E;
                // plane strain: lambda = v E / ((1 + v)(1 - 2 v)); plane stress: lambda' = v E / (1 - v^2)
                const double u2 = 1.0 - nu * nu;
                const double dl_dnu = M.stype == 1 ? E * (1.0 + nu * nu) / (u2 * u2) : E * (1.0 + 2.0 * nu * nu) / (t * t);
                const double dm_dnu = -0.5 * E / ((1.0 + nu) * (1.0 + nu));
                const double gE = gl * dl_dE + gm * dm_dE;
                const double gnu = gl * dl_dnu + gm * dm_dnu;
                // dE/dx0 = std0 * E ; dnu/dx1 = std1 * nu (1 - 2 nu)
                double *gxo = A.gx + (A.gx_stride ? (size_t)s * A.gx_stride + A.gx_off : (size_t)2 * s);
                gxo[0] = gE * M.theta_std[0] * E;
                gxo[1] = gnu * M.theta_std[1] * nu * (1.0 - 2.0 * nu);
            }
        }
        if (tid == 0 && A.status) A.status[s] = s_flag;
        __syncthreads();
    }
}

// Step-1 ELBO reductions over the local sample range (deterministic: fixed order per output).
//   sums[0..1] = sum_j f_j, sums[2] = sum_j |f_j|^2
//   gmu[b][k]  = sum_s gtheta[b,s][k],  gsig2[b][k] = sum_s gtheta[b,s][k] * e[s][k] / (2 sqrt(sig2[b][k]))
// PEER: the partial sums go straight into every rank's mailbox (vbfem_peer.cuh) and the last block to arrive
// writes the world's totals [sums(3) | gmu(2B) | gsig2(2B)] to `sums` (then one contiguous buffer).
template <bool PEER>
__global__ void elbo_reduce_kernel(int B, int S, long long j_begin, long long j_end, const double *f,
                                   const double *gth, const double *e, const double *sig2, double *sums,
                                   double *gmu, double *gsig2, const PeerCtx P) {
    __shared__ double sh[3][256];
    const int tid = threadIdx.x;
    unsigned long long q = 0;
    if (PEER) {
        q = *P.seq + 1;
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (long long j = j_begin + tid; j < j_end; j += blockDim.x) {
            const double f0 = f[2 * (j - j_begin)], f1 = f[2 * (j - j_begin) + 1];
            a0 += f0;
            a1 += f1;
            a2 += f0 * f0 + f1 * f1;
        }
        sh[0][tid] = a0;
        sh[1][tid] = a1;
        sh[2][tid] = a2;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (tid < o)
                for (int q = 0; q < 3; ++q) sh[q][tid] += sh[q][tid + o];
            __syncthreads();
        }
        if (tid < 3) {
            if (PEER) peer_push(P, q, tid, sh[tid][0]);
            else sums[tid] = sh[tid][0];
        }
    } else {
        // one block per observation row b: its (up to S) samples of the local range, strided over the threads, then a
        // fixed-order tree -- a rank that owns few rows with many samples each (S = 128 x GPUs) is as fast as one GPU
        const int bb = blockIdx.x - 1;
        long long lo = (long long)bb * S, hi = lo + S;
        lo = lo > j_begin ? lo : j_begin;
        hi = hi < j_end ? hi : j_end;
        double gm0 = 0.0, gm1 = 0.0, gs0 = 0.0, gs1 = 0.0;
        for (long long j = lo + tid; j < hi; j += blockDim.x) {
            const double2 g = *reinterpret_cast<const double2 *>(gth + 2 * (j - j_begin));
            const double2 ee = *reinterpret_cast<const double2 *>(e + 2 * (j - (long long)bb * S));
            gm0 += g.x;
            gm1 += g.y;
            gs0 = fma(g.x, ee.x, gs0);
            gs1 = fma(g.y, ee.y, gs1);
        }
        __shared__ double sh4[4][256];
        sh4[0][tid] = gm0;
        sh4[1][tid] = gm1;
        sh4[2][tid] = gs0;
        sh4[3][tid] = gs1;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (tid < o)
                for (int q4 = 0; q4 < 4; ++q4) sh4[q4][tid] += sh4[q4][tid + o];
            __syncthreads();
        }
        if (tid < 2) {
            const int idx = 2 * bb + tid;
            const double gm = sh4[tid][0];
            const double gs2 = sh4[2 + tid][0] * 0.5 / sqrt(sig2[idx]);
            if (PEER) {
                peer_push(P, q, 3 + idx, gm);
                peer_push(P, q, 3 + 2 * B + idx, gs2);
            } else {
                gmu[idx] = gm;
                gsig2[idx] = gs2;
            }
        }
    }
    if (PEER) peer_finish(P, q, (int)gridDim.x, 3 + 4 * B, sums);
}

// Step-2 sufficient statistics over the local sample range (fixed order):
//   sums[0..1] = sum_j h_j (per component), sums[2..3] = sum_j h_j^2
template <bool PEER>
__global__ void hsum_kernel(long long nloc, const double *h, double *sums, const PeerCtx P) {
    __shared__ double sh[4][256];
    const int tid = threadIdx.x;
    unsigned long long q = 0;
    if (PEER) {
        q = *P.seq + 1;
        __syncthreads();
    }
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (long long j = tid; j < nloc; j += blockDim.x) {
        const double h0 = h[2 * j], h1 = h[2 * j + 1];
        a0 += h0;
        a1 += h1;
        a2 += h0 * h0;
        a3 += h1 * h1;
    }
    sh[0][tid] = a0;
    sh[1][tid] = a1;
    sh[2][tid] = a2;
    sh[3][tid] = a3;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o)
            for (int q = 0; q < 4; ++q) sh[q][tid] += sh[q][tid + o];
        __syncthreads();
    }
    if (tid < 4) {
        if (PEER) peer_push(P, q, tid, sh[tid][0]);
        else sums[tid] = sh[tid][0];
    }
    if (PEER) peer_finish(P, q, 1, 4, sums);
}

__global__ void ysum_kernel(int B, const double *y, double *out) {
    if (threadIdx.x < 2) {
        double a = 0.0;
        for (int i = 0; i < B; ++i) a += y[2 * i + threadIdx.x];
        out[threadIdx.x] = a;
    } else if (threadIdx.x == 2) {  // sum_b |y_b|^2 (the loss value of term2 needs it)
        double a = 0.0;
        for (int i = 0; i < 2 * B; ++i) a += y[i] * y[i];
        out[2] = a;
    }
}

// The step-1 loss and its gradient w.r.t. the nets' outputs from the (all-reduced) totals of the data term:
//   loss = term1 - term2 - term3  (main_custom_training.py:183-185, 199-214, 226-235), theta_dim = y_dim = 2,
//   out[0] = loss, out[1 + 2b + k] = d loss / d mu[b][k], out[1 + 2B + ...] = d loss / d sig2, out[1 + 4B + ...] =
//   d loss / d log_sig2 (the direct dependence through term1 only; sig2 = exp(log_sig2) is chained by the caller).
__global__ void elbo_loss_kernel(int B, double nsamp, double sig_e, const double *tot, const double *mu,
                                 const double *sig2, const double *lsg, const double *ysum, double *out) {
    const int idx = threadIdx.x;
    const double invB = 1.0 / (double)B;
    if (idx < 2 * B) {
        out[1 + idx] = tot[3 + idx] + mu[idx] * invB;            // -term3 -> + mu / B
        out[1 + 2 * B + idx] = tot[3 + 2 * B + idx] + 0.5 * invB;  // -term3 -> + 1 / (2 B)
        out[1 + 4 * B + idx] = -0.5 * invB;                       // term1
    }
    if (idx == 0) {
        const double log2pi = 1.8378770664093454835606594728112;
        double sls = 0.0, s3 = 0.0;
        for (int i = 0; i < 2 * B; ++i) {
            sls += lsg[i];
            s3 += sig2[i] + mu[i] * mu[i];
        }
        const double t1 = -0.5 * sls * invB - 0.5 * 2.0 * log2pi - 0.5 * 2.0;
        const double t3 = -0.5 * 2.0 * log2pi - 0.5 * s3 * invB;
        const double total = (double)B * tot[2] - 2.0 * (tot[0] * ysum[0] + tot[1] * ysum[1]) + nsamp * ysum[2];
        const double t2 = -0.5 * 2.0 * log(2.0 * 3.14159265358979323846 * sig_e) - 0.5 / sig_e * total / ((double)B * nsamp);
        out[0] = t1 - t2 - t3;
    }
}

// Peak probes for the roofline denominators.
__global__ void dfma_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void copy_kernel(const double4 *__restrict__ src, double4 *__restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace vbfem

// ==========================================================================================
// Host side
// ==========================================================================================
using namespace vbfem;

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                        \
    } while (0)

typedef void (*kernel_fn)(const DevModel, const Args);
typedef void (*panel_fn)(const DevModel, const PanelModel, const Args);
typedef void (*warp_fn)(const DevModel, const WarpModel, const Args);

struct vbfem_handle {
    int device = 0;
    DevModel M{};
    // host copies used by the launcher
    int num_sms = 0, ctas_per_sm = 0, block = 0;
    size_t smem_bytes = 0;
    kernel_fn kern = nullptr;
    kernel_fn kern_front[3] = {nullptr, nullptr, nullptr};  // forward / forward+adjoint / Jacobian
    // blocked panel kernel (wide bands, factor streamed to HBM)
    PanelModel PM{};
    panel_fn kern_panel[3] = {nullptr, nullptr, nullptr};
    // warp-per-sample kernel (narrow bands, window in registers)
    WarpModel WM{};
    warp_fn kern_warp[3] = {nullptr, nullptr, nullptr};
    int warp_nw = 0;
    // second generation, forward / fused-adjoint modes: sixteen warps per SM when registers (128) and shared memory allow
    warp_fn kern_warp16[3] = {nullptr, nullptr, nullptr};
    size_t warp16_smem = 0, warp16j_smem = 0;   // j: Jacobian mode (five vectors in the window, the smallest ring)
    int warp16_per_warp = 0, warp16j_per_warp = 0, warp16j_rows = 0;
    // generic kernel configuration (fields mode, meshes neither fast kernel takes)
    DevModel M_gen{};
    int gen_block = 0, gen_ctas = 0;
    size_t gen_smem_bytes = 0;
    long long gen_ws_stride = 0;
    double *ws_gen = nullptr;  // per-CTA band scratch when the band does not fit in shared memory (sized at create)
    std::vector<void *> dev_allocs;
    // per-sample buffers: they grow in vbfem_reserve or, outside stream capture, on demand
    double *jac = nullptr;  // 4x2 Jacobians d(y, h)/dx kept by vbfem_forward(keep_factor=1)
    long long jac_cap = 0, kept_n = 0;
    int64_t ticket = 0, kept_ticket = 0;
    int *status = nullptr;
    long long status_cap = 0, last_n = 0;
    // ELBO scratch
    double *elbo_f = nullptr, *elbo_g = nullptr, *elbo_ysum = nullptr;
    long long elbo_cap = 0;
    // host entry points: pinned (mapped) staging, device staging, their own stream
    double *pin = nullptr, *pin_dev = nullptr, *stage = nullptr;
    long long stage_cap = 0;
    cudaStream_t host_stream = nullptr;
    int info_colors = 0;
    int n_real = 0;   // order of the system without padding rows
    int variant = 0;  // 0 = generic per-column kernel, 2 = on-chip front kernel, 3 = blocked panel kernel
    long long *timeline = nullptr;
    // NVLink peer mailboxes of the ELBO all-reduce (vbfem_peer.cuh)
    PeerCtx peer{};
    void *peer_mail = nullptr, *peer_state = nullptr;
    std::vector<void *> peer_opened;
    bool peer_ready = false;
};

template <typename T>
static int upload(vbfem_handle *h, const std::vector<T> &v, const T **out) {
    void *p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
    CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    h->dev_allocs.push_back(p);
    *out = (const T *)p;
    return 0;
}

// Half bandwidth over free dofs for a given node visiting order.
static int band_for_order(const vbfem_mesh *m, const std::vector<int> &order, const std::vector<char> &is_free,
                          std::vector<int> &dof2band) {
    const int ndof = 2 * m->nnodes;
    dof2band.assign(ndof, -1);
    int next = 0;
    for (int nd : order)
        for (int c = 0; c < 2; ++c)
            if (is_free[2 * nd + c]) dof2band[2 * nd + c] = next++;
    int bw = 0;
    for (int e = 0; e < m->nele; ++e) {
        int lo = 1 << 30, hi = -1;
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 2; ++c) {
                const int p = dof2band[2 * (m->ien[4 * e + a] - 1) + c];
                if (p >= 0) {
                    lo = std::min(lo, p);
                    hi = std::max(hi, p);
                }
            }
        if (hi >= 0) bw = std::max(bw, hi - lo);
    }
    return bw;
}

static std::vector<int> rcm_order(const vbfem_mesh *m) {
    const int nn = m->nnodes;
    std::vector<std::vector<int>> adj(nn);
    for (int e = 0; e < m->nele; ++e)
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 4; ++c)
                if (a != c) adj[m->ien[4 * e + a] - 1].push_back(m->ien[4 * e + c] - 1);
    for (auto &v : adj) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    std::vector<int> order;
    std::vector<char> seen(nn, 0);
    auto bfs = [&](int start, std::vector<int> &out, std::vector<char> &mark) {
        std::queue<int> q;
        q.push(start);
        mark[start] = 1;
        while (!q.empty()) {
            int u = q.front();
            q.pop();
            out.push_back(u);
            std::vector<int> nb;
            for (int v : adj[u])
                if (!mark[v]) {
                    mark[v] = 1;
                    nb.push_back(v);
                }
            std::sort(nb.begin(), nb.end(), [&](int a, int c) {
                return adj[a].size() != adj[c].size() ? adj[a].size() < adj[c].size() : a < c;
            });
            for (int v : nb) q.push(v);
        }
    };
    for (int root = 0; root < nn; ++root) {
        if (seen[root]) continue;
        int start = root;
        for (int it = 0; it < 3; ++it) {  // pseudo-peripheral node
            std::vector<int> tmp;
            std::vector<char> mk(nn, 0);
            bfs(start, tmp, mk);
            start = tmp.back();
        }
        bfs(start, order, seen);
    }
    std::reverse(order.begin(), order.end());
    return order;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel FUNCTION, not to a handle: two engines (two
// meshes) share it, and the one created last would cap the other's launches.  Always opt in to the device maximum.
static int g_smem_optin = 232448;
template <typename F>
static cudaError_t allow_max_smem(F k) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin - (int)fa.sharedSizeBytes);
}

template <int NT, int EPT, int MINB, bool HBM>
static int configure(vbfem_handle *h, size_t smem) {
    kernel_fn k = fem_kernel<NT, EPT, MINB, HBM>;
    if (smem > (size_t)g_smem_optin) return fail(-3, "kernel does not fit: %zu bytes of shared memory", smem);
    CU(allow_max_smem(k));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, NT, smem));
    if (nb < 1) return fail(-3, "kernel does not fit: %zu bytes of shared memory", smem);
    h->kern = k;
    h->block = NT;
    h->ctas_per_sm = nb;
    h->smem_bytes = smem;
    return 0;
}

// Host copy of shapef_q4 (vbfem_math.cuh) for the sample-independent observation geometry.
static void host_shapef_q4(const double *x, const double *y, int gp, double *nx, double *ny) {
    const double g = 0.577350269189626;
    const double s0 = (gp == 1 || gp == 2) ? g : -g, s1 = (gp >= 2) ? g : -g;
    const double sh = 0.5 * s0, th = 0.5 * s1;
    const double sp = 0.5 + sh, tp = 0.5 + th, sm = 0.5 - sh, tm = 0.5 - th;
    const double xo = x[0] - x[1] + x[2] - x[3];
    double xs = -x[0] + x[1] + x[2] - x[3] + xo * s1;
    double xt = -x[0] - x[1] + x[2] + x[3] + xo * s0;
    const double yo = y[0] - y[1] + y[2] - y[3];
    double ys = -y[0] + y[1] + y[2] - y[3] + yo * s1;
    double yt = -y[0] - y[1] + y[2] + y[3] + yo * s0;
    double xsj1 = xs * yt - xt * ys;
    xsj1 = (xsj1 != 0.0) ? 1.0 / xsj1 : 1.0;
    xs = (xs + xs) * xsj1;
    xt = (xt + xt) * xsj1;
    ys = (ys + ys) * xsj1;
    yt = (yt + yt) * xsj1;
    const double ytm = yt * tm, ysm = ys * sm, ytp = yt * tp, ysp = ys * sp;
    const double xtm = xt * tm, xsm = xs * sm, xtp = xt * tp, xsp = xs * sp;
    nx[0] = -ytm + ysm; nx[1] = ytm + ysp; nx[2] = ytp - ysp; nx[3] = -ytp - ysm;
    ny[0] = xtm - xsm; ny[1] = -xtm - xsp; ny[2] = -xtp + xsp; ny[3] = xtp + xsm;
}

// Internal numbering: try natural / coordinate-sorted / RCM node orders, keep the narrowest band.
// Returns the half bandwidth over the free dofs; dof2band[global dof] = band row or -1.
static int choose_numbering(const vbfem_mesh *m, const std::vector<char> &is_free, std::vector<int> &dof2band) {
    const int nn = m->nnodes;
    std::vector<std::vector<int>> cands;
    std::vector<int> nat(nn);
    std::iota(nat.begin(), nat.end(), 0);
    cands.push_back(nat);
    double ext = 0;
    for (int i = 0; i < 2 * nn; ++i) ext = std::max(ext, std::fabs(m->coord[i]));
    const double tol = std::max(ext, 1.0) * 1e-9;
    for (int major = 0; major < 2; ++major) {
        std::vector<int> o = nat;
        std::stable_sort(o.begin(), o.end(), [&](int a, int c) {
            const double pa = m->coord[2 * a + major], pc = m->coord[2 * c + major];
            if (std::fabs(pa - pc) > tol) return pa < pc;
            return m->coord[2 * a + 1 - major] < m->coord[2 * c + 1 - major];
        });
        cands.push_back(o);
    }
    cands.push_back(rcm_order(m));
    int best_bw = 1 << 30;
    std::vector<int> tmp;
    for (size_t c = 0; c < cands.size(); ++c) {
        const int bw = band_for_order(m, cands[c], is_free, tmp);
        if (bw < best_bw) {
            best_bw = bw;
            dof2band = tmp;
        }
    }
    return best_bw;
}

// Layout of the on-chip two-front kernel (vbfem_front_kernel.cuh) for a numbering: the band order is
// oriented so that it ENDS at the observed node (its unit vectors then ride along the bottom front), the
// observed element's dofs must fit into the P = 26 middle rows [pT, pT+P), the bottom front owns the
// last nB rows mirrored.  ok == false: the generic kernel serves this mesh / observation set-up.
struct FrontPlan {
    bool ok, flip;
    int pT;
    int tip[2], er[8];  // band rows (before orientation) of the observed node's / element's dofs
};
constexpr int kFrontB = 25, kFrontP = kFrontB + 1, kFrontNT = 128;
static size_t front_smem_bytes(int n) {
    return ((size_t)n * kFrontP + 5 * (size_t)n + 32 + 32 + 8 * (kFrontNT / 32)) * sizeof(double);
}
static FrontPlan plan_front(const vbfem_mesh *m, const std::vector<int> &dof2band, int n, int b,
                            size_t smem_per_sm) {
    constexpr int TB = kFrontB, TP = kFrontP;
    FrontPlan P{};
    int elo = 1 << 30, ehi = -1, tlo = 1 << 30, thi = -1;
    for (int k = 0; k < 2; ++k) {
        P.tip[k] = dof2band[2 * (m->obs_node - 1) + k];
        if (P.tip[k] >= 0) {
            tlo = std::min(tlo, P.tip[k]);
            thi = std::max(thi, P.tip[k]);
        }
    }
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 2; ++c) {
            const int g = dof2band[2 * (m->ien[4 * (m->obs_ele - 1) + a] - 1) + c];
            P.er[2 * a + c] = g;
            if (g >= 0) {
                elo = std::min(elo, g);
                ehi = std::max(ehi, g);
            }
        }
    bool ok = getenv("VBFEM_FORCE_GENERIC") == nullptr && b <= TB && ehi >= 0 && ehi - elo < TP;
    bool flip = false;
    if (ok && thi >= 0) {
        if (thi < elo)
            flip = true;  // observed node ahead of the observed element: reverse the band order
        else if (tlo <= ehi)
            ok = false;   // observed node inside the element's row range: generic kernel
    }
    int pT = 0;
    if (ok) {
        const int lo = flip ? n - 1 - ehi : elo, hi = flip ? n - 1 - elo : ehi;
        const int tmin = (thi < 0) ? n : (flip ? n - 1 - thi : tlo);
        // pT in [hi-P+1, lo], the observed node inside the bottom front, both fronts at least 32 columns
        const int pmin = std::max(hi - TP + 1, 32), pmax = std::min({lo, n - TP - 32, tmin - TP});
        if (pmin > pmax) ok = false;
        // balance: the top front also eliminates the middle block
        pT = std::min(std::max((int)((n - 2 * TP) * 0.565), pmin), pmax);
    }
    // the first two vectors double as the zero extension behind the bottom front's band
    // ((TP-1)*TP + 1 doubles) and as the Schur hand-over scratch (TP*TP + 3*TP + 4 doubles)
    if (ok && 2 * n < TP * TP + 3 * TP + 4) ok = false;
    if (ok && (n % 2 != 0 || (size_t)n * TP >= 32000 || 2 * (front_smem_bytes(n) + 1024) > smem_per_sm)) ok = false;
    P.ok = ok;
    P.flip = flip;
    P.pT = pT;
    return P;
}

// ------------------------------------------------------------------------------------------
// Blocked panel kernel (vbfem_panel.cuh): host-side plan.  Everything the kernel indexes by --
// padded band rows, block rows, the gather table that assembles a fresh block row from the ring of
// element matrices, the order in which the element matrices are needed, the initial right-hand-side
// blocks, the update schedule and the shared-memory layout -- is computed here, GPU-free.
// ------------------------------------------------------------------------------------------
struct PanelPlan {
    bool ok = false, flip = false;
    std::string why;
    int n = 0, off = 0, npad = 0, NQ = 0, NB = 0, R = 0, nub = 0, nele = 0;
    int obs_loc[2] = {-1, -1};
    int kstart[kPanelNW + 1] = {0};
    int o_win = 0, o_rhs = 0, o_lst = 0, o_ke = 0, o_lneg = 0, smem_bytes = 0, stages = 0;
    // second generation (window in registers; laid out for NB = kPanel2NB)
    bool v2 = false;
    int o2_wdiag = 0, o2_fresh = 0, o2_vst = 0, o2_big = 0, o2_lneg = 0, o2_rec = 0, o2_lst = 0, o2_ke = 0, smem2_bytes = 0,
        stages2 = 0;
    std::vector<int> gptr, eneed, eord, elm;
    std::vector<double> ecoord;
    std::vector<unsigned char> rec;  // row records (PanelModel::rec)
    int rec_stride = 0, rec_o_src = 0, rec_o_dst = 0, o_rec = 0;
    std::vector<unsigned short> gdst, gsrc, ub;  // gsrc: four entries per target
    std::vector<double> rhs0;
};

// nb_min > 0: the window is at least nb_min blocks wide (the warp kernel's register window has a fixed shape);
// jit_batch > 0: element matrices are computed `jit_batch` at a time straight into the ring, right before the
// first row that needs them (the warp kernel), instead of being fetched kPanelRecDepth rows ahead.
static PanelPlan plan_panel(const vbfem_mesh *m, const std::vector<int> &dof2band, int n, int nb_min = 0,
                            int jit_batch = 0) {
    PanelPlan P;
    auto no = [&](const char *why) {
        P.ok = false;
        P.why = why;
        return P;
    };
    const int ne = m->nele;
    P.n = n;
    P.nele = ne;
    if (n < 16) return no("fewer than 16 unknowns");
    // orientation: the free dofs of the observed node must be the LAST rows of the band (their unit
    // vectors then live in the last diagonal block)
    int tip[2], nt = 0, tlo = 1 << 30, thi = -1;
    for (int k = 0; k < 2; ++k) {
        tip[k] = dof2band[2 * (m->obs_node - 1) + k];
        if (tip[k] >= 0) {
            ++nt;
            tlo = std::min(tlo, tip[k]);
            thi = std::max(thi, tip[k]);
        }
    }
    if (nt > 0) {
        if (tlo == n - nt && thi == n - 1)
            P.flip = false;
        else if (tlo == 0 && thi == nt - 1)
            P.flip = true;
        else
            return no("the observed node is not at an end of the band order");
    }
    P.off = (8 - n % 8) % 8;
    P.npad = n + P.off;
    P.NQ = P.npad / 8;
    auto prow = [&](int r) { return r < 0 ? -1 : (P.flip ? n - 1 - r : r) + P.off; };
    for (int k = 0; k < 2; ++k) P.obs_loc[k] = tip[k] >= 0 ? prow(tip[k]) - 8 * (P.NQ - 1) : -1;

    // element rows, block half bandwidth, first / last block row of every element
    P.elm.assign((size_t)8 * ne, -1);
    std::vector<int> first(ne, 1 << 30), last(ne, -1);
    int NB = 1;
    for (int e = 0; e < ne; ++e) {
        int lo = 1 << 30, hi = -1;
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 2; ++c) {
                const int r = prow(dof2band[2 * (m->ien[4 * e + a] - 1) + c]);
                P.elm[8 * e + 2 * a + c] = r;
                if (r >= 0) {
                    lo = std::min(lo, r);
                    hi = std::max(hi, r);
                }
            }
        if (hi >= 0) {
            first[e] = lo / 8;
            last[e] = hi / 8;
            NB = std::max(NB, hi / 8 - lo / 8);
        }
    }
    NB = std::max(NB, nb_min);
    if (NB > kPanelNBMax) return no("band wider than the panel kernel's window");
    P.NB = NB;
    const int NB1 = NB + 1;

    // first-use order of the elements, how many of them block row q needs, ring capacity
    P.eord.resize(ne);
    std::iota(P.eord.begin(), P.eord.end(), 0);
    std::stable_sort(P.eord.begin(), P.eord.end(), [&](int a, int c) { return first[a] < first[c]; });
    std::vector<int> pos(ne);
    for (int k = 0; k < ne; ++k) pos[P.eord[k]] = k;
    P.ecoord.resize((size_t)8 * ne);
    for (int k = 0; k < ne; ++k)
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 2; ++c) P.ecoord[8 * k + 2 * a + c] = m->coord[2 * (m->ien[4 * P.eord[k] + a] - 1) + c];
    P.eneed.assign(P.NQ, 0);
    {
        int k = 0;
        for (int q = 0; q < P.NQ; ++q) {
            while (k < ne && first[P.eord[k]] <= q) ++k;
            P.eneed[q] = k;
        }
    }
    {
        std::vector<int> minpos(P.NQ + 1, 1 << 30);  // smallest position among elements still needed at row >= q
        for (int e = 0; e < ne; ++e)
            if (last[e] >= 0) minpos[last[e]] = std::min(minpos[last[e]], pos[e]);
        for (int q = P.NQ - 1; q >= 0; --q) minpos[q] = std::min(minpos[q], minpos[q + 1]);
        // the element matrices of row q are fetched into the ring kPanelRecDepth rows ahead of the gather
        int R = P.eneed[std::min(NB, P.NQ - 1)];
        if (jit_batch > 0) {
            R = 1;
            int computed = 0;
            for (int q = 0; q < P.NQ; ++q) {
                while (computed < P.eneed[q]) computed = std::min(computed + jit_batch, ne);
                if (minpos[q] < computed) R = std::max(R, computed - minpos[q]);
            }
        } else {
            for (int q = 0; q < P.NQ; ++q) {
                const int ahead = P.eneed[std::min(q + kPanelRecDepth, P.NQ - 1)];
                if (minpos[q] < ahead) R = std::max(R, ahead - minpos[q]);
            }
        }
        P.R = R;
    }
    if ((size_t)P.R * 36 + 2 > 65535) return no("element ring too large for 16-bit gather indices");
    const unsigned short ZERO = (unsigned short)(P.R * 36), ONE = (unsigned short)(P.R * 36 + 1);

    // gather table: target (block row, d * 64 + g * 8 + c) <- element-ring entries
    struct Tgt {
        int n = 0;
        unsigned short s[4];
    };
    std::vector<std::map<int, Tgt>> rows(P.NQ);
    for (int r = 0; r < P.off; ++r) {  // leading pad rows: identity
        Tgt &t = rows[0][r * 8 + r];
        t.n = 1;
        t.s[0] = ONE;
    }
    for (int e = 0; e < ne; ++e) {
        const int slot = pos[e] % P.R;
        for (int a = 0; a < 8; ++a)
            for (int q = 0; q <= a; ++q) {
                const int ra = P.elm[8 * e + a], rq = P.elm[8 * e + q];
                if (ra < 0 || rq < 0) continue;
                const int rr = std::max(ra, rq), cc = std::min(ra, rq);
                const int d = rr / 8 - cc / 8;
                Tgt &t = rows[rr / 8][d * 64 + (rr % 8) * 8 + cc % 8];
                if (t.n >= 4) return no("more than four elements share a matrix entry");
                t.s[t.n++] = (unsigned short)(slot * 36 + tri(a, q));
            }
    }
    P.gptr.assign(P.NQ + 1, 0);
    for (int q = 0; q < P.NQ; ++q) {
        for (auto &kv : rows[q]) {
            P.gdst.push_back((unsigned short)kv.first);
            for (int i = 0; i < 4; ++i) P.gsrc.push_back(i < kv.second.n ? kv.second.s[i] : ZERO);
        }
        P.gptr[q + 1] = (int)P.gdst.size();
    }

    // initial right-hand-side blocks [q][a][c]: row 0 the load vector, rows 1..6 the strain functionals
    // (exx, eyy, gxy) of the two observed Gauss points (src/mat_subroutine_tf.py:112-145 as vectors)
    P.rhs0.assign((size_t)P.NQ * 64, 0.0);
    for (int i = 0; i < n; ++i) {
        const int r = prow(dof2band[m->free_dof[i] - 1]);
        P.rhs0[(size_t)(r / 8) * 64 + r % 8] = m->pf[i];
    }
    {
        double ox[4], oy[4];
        int orow[8];
        for (int a = 0; a < 4; ++a) {
            const int nd = m->ien[4 * (m->obs_ele - 1) + a] - 1;
            ox[a] = m->coord[2 * nd];
            oy[a] = m->coord[2 * nd + 1];
            for (int c = 0; c < 2; ++c) orow[2 * a + c] = prow(dof2band[2 * nd + c]);
        }
        for (int g = 0; g < 2; ++g) {
            double nx[4], ny[4];
            host_shapef_q4(ox, oy, m->obs_gp[g] - 1, nx, ny);
            auto add = [&](int row, int r, double v) {
                if (r >= 0) P.rhs0[(size_t)(r / 8) * 64 + row * 8 + r % 8] += v;
            };
            for (int a = 0; a < 4; ++a) {
                add(1 + 3 * g, orow[2 * a], nx[a]);      // exx = sum dN/dx ux
                add(2 + 3 * g, orow[2 * a + 1], ny[a]);  // eyy = sum dN/dy uy
                add(3 + 3 * g, orow[2 * a + 1], nx[a]);  // gxy = sum dN/dx uy + dN/dy ux
                add(3 + 3 * g, orow[2 * a], ny[a]);
            }
        }
    }

    // trailing-update schedule of warps 0..5: blocks (I, J), 1 <= J <= I <= NB, then the right-hand-side row
    // I = NB+1; block (1, 1) -- the next diagonal block -- belongs to the look-ahead warp
    for (int I = 1; I <= NB + 1; ++I)
        for (int J = 1; J <= std::min(I, NB); ++J)
            if (!(I == 1 && J == 1)) P.ub.push_back((unsigned short)((I << 8) | J));
    P.nub = (int)P.ub.size();
    for (int w = 0; w <= kPanelNW; ++w) P.kstart[w] = (int)((long long)P.nub * std::min(w, kPanelUpdW) / kPanelUpdW);

    // row records
    {
        int maxnew = 0, maxcnt = 0;
        for (int q = 0; q < P.NQ; ++q) {
            maxnew = std::max(maxnew, P.eneed[q] - (q ? P.eneed[q - 1] : 0));
            maxcnt = std::max(maxcnt, P.gptr[q + 1] - P.gptr[q]);
        }
        (void)maxnew;
        P.rec_o_src = 16 + 512;
        P.rec_o_dst = P.rec_o_src + 8 * maxcnt;
        P.rec_stride = (P.rec_o_dst + 2 * maxcnt + 15) & ~15;
        P.rec.assign((size_t)P.NQ * P.rec_stride, 0);
        for (int q = 0; q < P.NQ; ++q) {
            unsigned char *rc = P.rec.data() + (size_t)q * P.rec_stride;
            const int e0 = q ? P.eneed[q - 1] : 0, nnew = P.eneed[q] - e0, cnt = P.gptr[q + 1] - P.gptr[q];
            int hd[4] = {nnew, cnt, e0, 0};
            memcpy(rc, hd, 16);
            memcpy(rc + 16, P.rhs0.data() + (size_t)q * 64, 512);
            memcpy(rc + P.rec_o_src, P.gsrc.data() + (size_t)4 * P.gptr[q], (size_t)8 * cnt);
            memcpy(rc + P.rec_o_dst, P.gdst.data() + P.gptr[q], (size_t)2 * cnt);
        }
    }

    // shared-memory layout: window (diagonal d: ring of NB+2-d blocks), rhs ring, two staging panels, element ring
    const int nwin = NB1 * (NB + 4) / 2, LPBb = (NB + 2) * 512;
    P.o_win = (int)((sizeof(PanelSmem) + 15) & ~(size_t)15);
    P.o_rhs = P.o_win + nwin * 512;
    P.o_lst = P.o_rhs + (NB + 2) * 512;
    P.o_ke = P.o_lst + 2 * LPBb;
    const int ke_bytes = std::max((P.R * 36 + 2) * 8, (NB1 + 2 * kPanelNW) * 512);
    P.o_lneg = P.o_ke + ((ke_bytes + 15) & ~15);
    P.o_rec = P.o_lneg + (NB + 3) * 512;
    P.smem_bytes = P.o_rec + kPanelRecDepth * P.rec_stride;
    P.stages = std::min(kPanelStagesMax, (nwin + NB + 2) * 512 / LPBb);
    if (P.stages < 2) return no("window too small for the reverse pass");
    P.ok = true;
    {   // second generation: shared memory without the window; the diagonal ring, the entering-row staging area, the V
        // blocks, -L and the record ring form one region that the reverse pass re-uses as its bulk-load ring
        int o = (int)((sizeof(PanelSmem) + 15) & ~(size_t)15);
        P.o2_big = o;
        P.o2_wdiag = o;
        o += (NB + 2) * 512;
        P.o2_fresh = o;
        o += (NB + 2) * 512;
        P.o2_vst = o;
        o += NB1 * 512;
        P.o2_lneg = o;
        o += (NB + 3) * 512;
        P.o2_rec = o;
        o += kPanelRecDepth * P.rec_stride;
        const int big = o - P.o2_big;
        P.o2_lst = o;
        o += 2 * LPBb;
        P.o2_ke = o;
        o += (ke_bytes + 15) & ~15;
        P.smem2_bytes = o;
        P.stages2 = std::min(kPanelStagesMax, big / LPBb);
        P.v2 = NB == kPanel2NB && P.stages2 >= 2 && P.NQ > NB1;
    }
    return P;
}

// Warp-per-sample kernel (vbfem_warp.cuh): chosen for narrow bands when enabled (VBFEM_WARP=0/1 overrides the default).
constexpr bool kWarpDefault = true;
static bool want_warp_kernel() {
    if (getenv("VBFEM_FORCE_GENERIC") != nullptr || getenv("VBFEM_FORCE_PANEL") != nullptr) return false;
    if (const char *e = getenv("VBFEM_WARP")) return atoi(e) != 0;
    return kWarpDefault;
}

static int warp_kernel_batch() {
    int b = 30;  // 30 instead of 32 elements per batch: ring of 44 instead of 46 matrices on Cook 20x10 -> 14 warps fit
    if (const char *e = getenv("VBFEM_WARP_BATCH")) b = atoi(e);
    return std::min(std::max(b, 1), kWarpBatch);
}
static int warp_kernel_smem(const PanelPlan &P) { return (kWarpFixed + (P.R * 36 + 2) * 8 + 15) & ~15; }
static int warp_kernel_tab_bytes(const PanelPlan &P) { return (int)((P.gdst.size() * 8 + (size_t)(P.NQ + 1) * 8 + 15) & ~(size_t)15); }
static bool warp_kernel_ok(const PanelPlan &P, int NW, size_t smem_per_block);
// Warps (= samples in flight) per CTA: twelve when the per-warp areas and the shared tables fit, else eight.
static int warp_kernel_warps(const PanelPlan &P, size_t smem_per_block) {
    if (const char *e = getenv("VBFEM_WARP_NW")) return atoi(e);
    return warp_kernel_ok(P, 12, smem_per_block) ? 12 : 8;
}
static bool warp_kernel_ok(const PanelPlan &P, int NW, size_t smem_per_block) {
    // 11-bit ring indices and 8-bit targets in the packed gather table
    return P.ok && P.NB == kWarpNB && P.NQ > kWarpNB && P.R * 36 + 2 <= 2048 && (NW == 8 || NW == 12 || NW == 16) &&
           (size_t)NW * warp_kernel_smem(P) + warp_kernel_tab_bytes(P) <= smem_per_block;
}

// Second generation of the warp kernel (vbfem_warp2.cuh): the (K_lam, K_mu) band table must fit in shared memory
// next to the per-warp areas of twelve (or eight) warps.
struct Warp2Plan {
    bool ok = false, ok16 = false, ok16j = false;  // sixteen warps per SM: forward / fused-adjoint mode, Jacobian mode
    int NW = 0, hb = 0, ldt = 0, tab_bytes = 0, warp_smem = 0, warp16_smem = 0, warp16j_smem = 0, rows16j = 0;
    size_t smem = 0, smem16 = 0, smem16j = 0;
};
static Warp2Plan warp2_plan(const PanelPlan &P, size_t smem_per_block) {
    Warp2Plan W;
    if (!P.ok || P.NB != kWarpNB || P.NQ <= kWarpNB || P.NQ > 128 || getenv("VBFEM_WARP_V1")) return W;
    int hb = 0;
    for (int e = 0; e < P.nele; ++e) {
        int lo = 1 << 30, hi = -1;
        for (int a = 0; a < 8; ++a) {
            const int r = P.elm[8 * e + a];
            if (r >= 0) {
                lo = std::min(lo, r);
                hi = std::max(hi, r);
            }
        }
        if (hi >= 0) hb = std::max(hb, hi - lo);
    }
    W.hb = hb;
    W.ldt = (hb + 2) & ~1;  // even row stride: rows hb + 1 apart land an odd number of 16-byte units apart
    W.tab_bytes = P.npad * W.ldt * 16;
    W.warp_smem = warp2_smem_per_warp(5);
    int nw = 0;
    if (const char *e = getenv("VBFEM_WARP_NW")) nw = atoi(e);
    for (int NW : {12, 8}) {
        if (nw && NW != nw) continue;
        if ((size_t)W.tab_bytes + (size_t)NW * W.warp_smem + 64 <= smem_per_block) {
            W.NW = NW;
            W.smem = (size_t)W.tab_bytes + (size_t)NW * W.warp_smem;
            W.ok = true;
            break;
        }
    }
    W.warp16_smem = warp2_smem_per_warp(2);
    W.smem16 = (size_t)W.tab_bytes + (size_t)16 * W.warp16_smem;
    const char *e16 = getenv("VBFEM_WARP2_NW16");
    W.ok16 = W.ok && W.NW == 12 && W.smem16 + 64 <= smem_per_block && !(e16 && atoi(e16) == 0) && !nw;
    W.rows16j = std::max(hb + 8, kWarp2WinRowsMin);
    W.warp16j_smem = warp2_smem_per_warp(5, W.rows16j);
    W.smem16j = (size_t)W.tab_bytes + (size_t)16 * W.warp16j_smem;
    W.ok16j = W.ok16 && W.rows16j <= kWarp2WinRows && W.smem16j + 64 <= smem_per_block;
    return W;
}

extern "C" const char *vbfem_last_error(void) { return g_err.c_str(); }

extern "C" int vbfem_create(vbfem_t **out, const vbfem_mesh *m, int device) {
    return vbfem_create_ex(out, m, nullptr, device);
}

extern "C" int vbfem_create_ex(vbfem_t **out, const vbfem_mesh *m, const vbfem_options *opt, int device) {
    if (!out || !m) return fail(-1, "null argument");
    const int stype = opt ? opt->stype : 2;
    if (stype != 1 && stype != 2) return fail(-1, "section stype %d not supported (1 plane stress, 2 plane strain)", stype);
    if (m->nnodes <= 0 || m->nele <= 0 || m->nfree <= 0 || !m->coord || !m->ien || !m->free_dof || !m->pf)
        return fail(-1, "incomplete mesh description");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(-4, "no CUDA device available (%s): libvbfem has no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(-1, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    const int nn = m->nnodes, ne = m->nele, ndof = 2 * nn;
    for (int i = 0; i < 4 * ne; ++i)
        if (m->ien[i] < 1 || m->ien[i] > nn) return fail(-1, "IEN entry %d out of range", m->ien[i]);
    if (m->obs_node < 1 || m->obs_node > nn || m->obs_ele < 1 || m->obs_ele > ne)
        return fail(-1, "observation node/element out of range");
    for (int k = 0; k < 2; ++k)
        if (m->obs_gp[k] < 1 || m->obs_gp[k] > 4) return fail(-1, "observation Gauss point out of range");
    std::vector<char> is_free(ndof, 0);
    for (int i = 0; i < m->nfree; ++i) {
        const int g = m->free_dof[i];
        if (g < 1 || g > ndof) return fail(-1, "free_dof entry %d out of range", g);
        is_free[g - 1] = 1;
    }

    // ---- internal numbering (narrowest band over natural / coordinate-sorted / RCM node orders)
    std::vector<int> dof2band;
    const int best_bw = choose_numbering(m, is_free, dof2band);
    const int n = m->nfree, b = std::max(best_bw, 1), ldb = b + 2;
    if (n > 32767) return fail(-3, "system too large for 16-bit band rows (n = %d)", n);

    // ---- element colouring (no two elements of a colour share a node) and colour-sorted order
    std::vector<int> color(ne, -1);
    std::vector<std::vector<int>> node_elems(nn);
    for (int e = 0; e < ne; ++e)
        for (int a = 0; a < 4; ++a) node_elems[m->ien[4 * e + a] - 1].push_back(e);
    int ncolors = 0;
    for (int e = 0; e < ne; ++e) {
        unsigned used = 0;
        for (int a = 0; a < 4; ++a)
            for (int o : node_elems[m->ien[4 * e + a] - 1])
                if (color[o] >= 0) used |= 1u << color[o];
        int c = 0;
        while (used & (1u << c)) ++c;
        if (c >= kMaxColors) return fail(-3, "mesh needs more than %d element colours", kMaxColors);
        color[e] = c;
        ncolors = std::max(ncolors, c + 1);
    }
    // Structured quad grids numbered along the short side (Cook nx x ny: b = 2(ny+2)+1) also admit the
    // colouring (i + 2j) mod 4.  It is preferred when valid: elements of one colour then spread over four
    // shared-memory bank classes instead of two (same-colour band targets lie 4 or 44 rows apart, a row
    // is 26 doubles), which halves the bank conflicts of the scatter.
    if (ncolors == 4 && b >= 9 && (b - 1) % 2 == 0) {
        const int per = 2 * ((b - 1) / 2 - 1);  // dofs per node column
        std::vector<int> alt(ne, -1);
        bool valid = per > 0;
        for (int e = 0; e < ne && valid; ++e) {
            int rmax = -1;
            for (int a = 0; a < 4; ++a)
                for (int c = 0; c < 2; ++c) rmax = std::max(rmax, dof2band[2 * (m->ien[4 * e + a] - 1) + c]);
            if (rmax < 0) valid = false;
            else alt[e] = ((rmax / per) + 2 * ((rmax % per) / 2)) & 3;
        }
        for (int e = 0; e < ne && valid; ++e)
            for (int a = 0; a < 4 && valid; ++a)
                for (int o : node_elems[m->ien[4 * e + a] - 1])
                    if (o != e && alt[o] == alt[e]) valid = false;
        if (valid) color = alt;
    }
    std::vector<int> eorder(ne);
    std::iota(eorder.begin(), eorder.end(), 0);
    std::stable_sort(eorder.begin(), eorder.end(), [&](int a, int c) { return color[a] < color[c]; });

    vbfem_handle *h = new vbfem_handle();
    struct Guard {  // every failing return below releases the half-built handle
        vbfem_handle *p;
        ~Guard() {
            if (p) vbfem_destroy(p);
        }
    } guard{h};
    h->device = device;
    DevModel &M = h->M;
    M.n = n;
    M.b = b;
    M.ldb = ldb;
    M.nele = ne;
    M.nnodes = nn;
    M.ndof = ndof;
    M.ncolors = ncolors;
    M.nitems = b * (b + 1) / 2 + b;
    M.thk = m->thk;
    M.stype = stype;
    for (int k = 0; k < 2; ++k) {
        M.theta_mean[k] = m->theta_mean[k];
        M.theta_std[k] = m->theta_std[k];
        M.obs_gp[k] = m->obs_gp[k] - 1;
        M.obs_dof[k] = dof2band[2 * (m->obs_node - 1) + k];
    }
    M.obs_ele = m->obs_ele - 1;
    M.w_first = n;
    for (int k = 0; k < 2; ++k)
        if (M.obs_dof[k] >= 0) M.w_first = std::min(M.w_first, M.obs_dof[k]);
    for (int a = 0; a < 4; ++a) {
        const int nd = m->ien[4 * M.obs_ele + a] - 1;
        M.obs_x[a] = m->coord[2 * nd];
        M.obs_y[a] = m->coord[2 * nd + 1];
        for (int c = 0; c < 2; ++c) {
            M.obs_lmb[2 * a + c] = dof2band[2 * nd + c];
            if (dof2band[2 * nd + c] >= 0) M.w_first = std::min(M.w_first, dof2band[2 * nd + c]);
        }
    }
    if (M.w_first >= n) M.w_first = 0;
    for (int c = 0; c <= kMaxColors; ++c) M.color_start[c] = ne;
    {
        int pos = 0;
        for (int c = 0; c < ncolors; ++c) {
            M.color_start[c] = pos;
            while (pos < ne && color[eorder[pos]] == c) ++pos;
        }
        M.color_start[ncolors] = ne;
    }

    // ---- device tables
    std::vector<double> coord(m->coord, m->coord + 2 * nn), pf(n, 0.0);
    std::vector<int> ien(4 * ne), lmg(8 * ne), band2dof(n);
    std::vector<short> lmb(8 * ne);
    for (int e = 0; e < ne; ++e)
        for (int a = 0; a < 4; ++a) {
            const int nd = m->ien[4 * e + a] - 1;
            ien[4 * e + a] = nd;
            for (int c = 0; c < 2; ++c) {
                lmg[8 * e + 2 * a + c] = 2 * nd + c;
                lmb[8 * e + 2 * a + c] = (short)dof2band[2 * nd + c];
            }
        }
    for (int i = 0; i < n; ++i) {
        const int g = m->free_dof[i] - 1;
        pf[dof2band[g]] = m->pf[i];
        band2dof[dof2band[g]] = g;
    }
    int rc = 0;
    rc |= upload(h, coord, &M.coord);
    rc |= upload(h, ien, &M.ien);
    rc |= upload(h, lmb, &M.lmb);
    rc |= upload(h, lmg, &M.lmg);
    rc |= upload(h, eorder, &M.eorder);
    rc |= upload(h, pf, &M.pf);
    rc |= upload(h, band2dof, &M.band2dof);
    if (rc) {
        return -2;
    }

    // ---- generic kernel configuration (any bandwidth; full fields): band in shared memory when it
    //      fits, else in HBM
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    g_smem_optin = (int)prop.sharedMemPerBlockOptin;
    {
        int NT = (M.nitems <= 352) ? 352 : (M.nitems <= 512 ? 256 : (M.nitems <= 4096 ? 512 : 1024));
        {   // band in HBM: 256 threads (255 registers, no spills: the per-column loop must not touch local memory)
            const size_t need = (((size_t)n * ldb + 1) & ~(size_t)1) * sizeof(double) +
                                (2 * (size_t)n + 2 * (NT / 32) + 32) * sizeof(double);
            if (need > (size_t)prop.sharedMemPerBlockOptin) NT = 256;
        }
        const int scratch = 2 * (NT / 32) + 32;
        const size_t band_doubles = ((size_t)n * ldb + 1) & ~(size_t)1;
        const size_t small = (2 * (size_t)n + scratch) * sizeof(double);
        const size_t big = small + band_doubles * sizeof(double);
        const size_t cap = (size_t)prop.sharedMemPerBlockOptin;
        M.band_in_smem = big <= cap ? 1 : 0;
        M.vec_off = M.band_in_smem ? (int)band_doubles : 0;
        M.red_off = M.vec_off + 2 * n;
        // band in HBM: ring of b + 1 + kRingDepth columns behind the vectors (factor window, sweep staging)
        M.ring_w = (b + 1 + kRingDepth + 1) & ~1;
        M.ring_off = (M.red_off + scratch + 1) & ~1;
        const size_t ring_bytes = M.band_in_smem ? 0 : (size_t)M.ring_w * ldb * sizeof(double);
        if (!M.band_in_smem && small + 16 + ring_bytes > cap) {
            return fail(-3, "mesh too large: %zu bytes of shared memory for the work vectors and the band window",
                        small + ring_bytes);
        }
        const size_t smem = M.band_in_smem ? big : small + 16 + ring_bytes;
        if (M.nitems > 1024 * 16) {
            return fail(-3, "half bandwidth %d too large", b);
        }
        if (!M.band_in_smem)
            rc = configure<256, 1, 1, true>(h, smem);
        else if (NT == 352)
            rc = configure<352, 1, 2, false>(h, smem);
        else if (NT == 256)
            rc = configure<256, 2, 2, false>(h, smem);
        else if (NT == 512)
            rc = configure<512, 8, 1, false>(h, smem);
        else
            rc = configure<1024, 16, 1, false>(h, smem);
        if (rc) {
            return rc;
        }
        h->gen_ws_stride = (long long)band_doubles + 8;
        h->gen_block = h->block;
        h->gen_ctas = h->ctas_per_sm;
        h->gen_smem_bytes = h->smem_bytes;
        h->M_gen = h->M;
    }
    CU(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&h->elbo_ysum, 8 * sizeof(double)));
    h->info_colors = ncolors;
    const bool force_panel = getenv("VBFEM_FORCE_PANEL") != nullptr;
    if (stype != 2) {  // plane stress: the generic kernel serves every mode (the fast kernels are plane strain only)
        guard.p = nullptr;
        *out = h;
        return 0;
    }

    // ---- warp-per-sample kernel: narrow bands (block half bandwidth <= 3), window in registers, 12 samples per SM
    if (!force_panel && want_warp_kernel()) {
        PanelPlan P = plan_panel(m, dof2band, n, kWarpNB, warp_kernel_batch());
        const int NW = warp_kernel_warps(P, (size_t)prop.sharedMemPerBlockOptin);
        const int warp_smem = warp_kernel_smem(P);
        const Warp2Plan W2 = warp2_plan(P, (size_t)prop.sharedMemPerBlockOptin);
        if (warp_kernel_ok(P, NW, (size_t)prop.sharedMemPerBlockOptin) && W2.ok) {
            // ---- second generation: K = lambda K_lam + mu K_mu from the band table in shared memory
            WarpModel &Q = h->WM;
            Q.n = P.n;
            Q.off = P.off;
            Q.npad = P.npad;
            Q.NQ = P.NQ;
            Q.R = 0;
            Q.nele = ne;
            Q.batch = 0;
            Q.obs_loc[0] = P.obs_loc[0];
            Q.obs_loc[1] = P.obs_loc[1];
            Q.warp_smem = W2.warp_smem;
            Q.tab_bytes = W2.tab_bytes;
            Q.ldt = W2.ldt;
            Q.hb = W2.hb;
            Q.win_rows = kWarp2WinRows;
            Q.cmagic = 65536u / (unsigned)(W2.hb + 1) + 1u;
            bool magic_ok = true;
            for (int id = 0; id < 8 * (W2.hb + 1); ++id)
                magic_ok &= (int)(((unsigned)id * Q.cmagic) >> 16) == id / (W2.hb + 1);
            Q.rhsmask[0] = Q.rhsmask[1] = 0;
            for (int q = 0; q < P.NQ; ++q)
                for (int i = 0; i < 64; ++i)
                    if (P.rhs0[(size_t)q * 64 + i] != 0.0) Q.rhsmask[q >> 6] |= 1ull << (q & 63);
            Q.rhs_first = P.NQ;
            for (int q = P.NQ - 1; q >= 0; --q)
                if ((Q.rhsmask[q >> 6] >> (q & 63)) & 1) Q.rhs_first = q;
            warp_fn ks[3];
            if (W2.NW == 8) {
                ks[0] = fem_warp2_kernel<0, 8>; ks[1] = fem_warp2_kernel<1, 8>; ks[2] = fem_warp2_kernel<2, 8>;
            } else {
                ks[0] = fem_warp2_kernel<0, 12>; ks[1] = fem_warp2_kernel<1, 12>; ks[2] = fem_warp2_kernel<2, 12>;
            }
            bool fits = magic_ok;
            for (int q = 0; q < 3 && fits; ++q) {
                cudaError_t e1 = allow_max_smem(ks[q]);
                int nb = 0;
                if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ks[q], W2.NW * 32, W2.smem);
                if (e1 != cudaSuccess || nb < 1) {
                    fits = false;
                    cudaGetLastError();
                }
                h->kern_warp[q] = ks[q];
            }
            if (fits) {
                // unit element matrices on the device (the same shapef_q4 / accumulate_kt as every kernel), summed on
                // the host in element order into the band table
                int rc2 = upload(h, P.ecoord, &Q.ecoord);
                if (rc2) return -2;
                double *dke = nullptr;
                CU(cudaMalloc(&dke, (size_t)ne * 72 * sizeof(double)));
                warp2_unit_element_kernel<<<(ne + 63) / 64, 64>>>(Q.ecoord, ne, m->thk, dke);
                std::vector<double> hke((size_t)ne * 72);
                cudaError_t ce = cudaMemcpy(hke.data(), dke, hke.size() * sizeof(double), cudaMemcpyDeviceToHost);
                cudaFree(dke);
                CU(ce);
                std::vector<double> tab((size_t)P.npad * W2.ldt * 2, 0.0);
                for (int r = 0; r < P.off; ++r) tab[((size_t)r * W2.ldt) * 2 + 1] = 1.0;  // pad rows: pivot mu, decoupled
                std::vector<int> pos(ne);
                for (int k = 0; k < ne; ++k) pos[P.eord[k]] = k;
                for (int e = 0; e < ne; ++e)
                    for (int a = 0; a < 8; ++a)
                        for (int q = 0; q <= a; ++q) {
                            const int ra = P.elm[8 * e + a], rq = P.elm[8 * e + q];
                            if (ra < 0 || rq < 0) continue;
                            const int rr = std::max(ra, rq), cc = std::min(ra, rq);
                            double *dst = &tab[((size_t)rr * W2.ldt + (rr - cc)) * 2];
                            dst[0] += hke[(size_t)pos[e] * 72 + tri(a, q)];
                            dst[1] += hke[(size_t)pos[e] * 72 + 36 + tri(a, q)];
                        }
                const double *dtab = nullptr;
                rc2 |= upload(h, tab, &dtab);
                Q.ktab = reinterpret_cast<const double2 *>(dtab);
                rc2 |= upload(h, P.rhs0, &Q.rhs0);
                if (rc2) return -2;
                long long nwarps = (long long)h->num_sms * W2.NW;
                {   // sixteen warps for MODE 0 / 1 (per-warp window of u and ONE adjoint vector)
                    const int pw = W2.warp16_smem;
                    const size_t sm16 = W2.smem16;
                    if (W2.ok16) {
                        warp_fn k16[2] = {fem_warp2_kernel<0, 16>, fem_warp2_kernel<1, 16>};
                        bool ok16 = true;
                        for (int q = 0; q < 2 && ok16; ++q) {
                            cudaError_t e1 = allow_max_smem(k16[q]);
                            int nb = 0;
                            if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k16[q], 512, sm16);
                            if (e1 != cudaSuccess || nb < 1) {
                                ok16 = false;
                                cudaGetLastError();
                            }
                        }
                        if (ok16) {
                            h->kern_warp16[0] = k16[0];
                            h->kern_warp16[1] = k16[1];
                            h->warp16_smem = sm16;
                            h->warp16_per_warp = pw;
                            nwarps = (long long)h->num_sms * 16;
                            if (W2.ok16j) {
                                warp_fn kj = fem_warp2_kernel<2, 16>;
                                cudaError_t e1 = allow_max_smem(kj);
                                int nb = 0;
                                if (e1 == cudaSuccess)
                                    e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kj, 512, W2.smem16j);
                                if (e1 == cudaSuccess && nb >= 1) {
                                    h->kern_warp16[2] = kj;
                                    h->warp16j_smem = W2.smem16j;
                                    h->warp16j_per_warp = W2.warp16j_smem;
                                    h->warp16j_rows = W2.rows16j;
                                } else {
                                    cudaGetLastError();
                                }
                            }
                        }
                    }
                }
                Q.lws_stride = (long long)P.NQ * (kWarpNB + 2) * 64;
                void *pl = nullptr;
                CU(cudaMalloc(&pl, (size_t)nwarps * Q.lws_stride * sizeof(double)));
                h->dev_allocs.push_back(pl);
                Q.lws = (double *)pl;
                h->warp_nw = W2.NW;
                h->block = W2.NW * 32;
                h->ctas_per_sm = 1;
                h->smem_bytes = W2.smem;
                h->variant = 4;
                h->n_real = n;
                h->PM.NB = kWarpNB;
                h->PM.R = 0;
                guard.p = nullptr;
                *out = h;
                return 0;
            }
        }
        if (warp_kernel_ok(P, NW, (size_t)prop.sharedMemPerBlockOptin)) {
            WarpModel &Q = h->WM;
            Q.n = P.n;
            Q.off = P.off;
            Q.npad = P.npad;
            Q.NQ = P.NQ;
            Q.R = P.R;
            Q.nele = ne;
            Q.batch = warp_kernel_batch();
            Q.obs_loc[0] = P.obs_loc[0];
            Q.obs_loc[1] = P.obs_loc[1];
            Q.warp_smem = warp_smem;
            warp_fn ks[3];
            if (NW == 8) {
                ks[0] = fem_warp_kernel<0, 8>; ks[1] = fem_warp_kernel<1, 8>; ks[2] = fem_warp_kernel<2, 8>;
            } else if (NW == 12) {
                ks[0] = fem_warp_kernel<0, 12>; ks[1] = fem_warp_kernel<1, 12>; ks[2] = fem_warp_kernel<2, 12>;
            } else {
                ks[0] = fem_warp_kernel<0, 16>; ks[1] = fem_warp_kernel<1, 16>; ks[2] = fem_warp_kernel<2, 16>;
            }
            bool fits = true;
            const size_t smem = (size_t)NW * warp_smem + warp_kernel_tab_bytes(P);
            for (int q = 0; q < 3 && fits; ++q) {
                cudaError_t e1 = allow_max_smem(ks[q]);
                int nb = 0;
                if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ks[q], NW * 32, smem);
                if (e1 != cudaSuccess || nb < 1) {
                    fits = false;
                    cudaGetLastError();
                }
                h->kern_warp[q] = ks[q];
            }
            if (fits) {
                std::vector<unsigned long long> gpack(P.gdst.size());
                for (size_t i = 0; i < P.gdst.size(); ++i) {
                    unsigned long long w = (unsigned long long)P.gdst[i] << 44;
                    for (int k = 0; k < 4; ++k) w |= (unsigned long long)P.gsrc[4 * i + k] << (11 * k);
                    gpack[i] = w;
                }
                std::vector<int2> rowtab(P.NQ + 1);
                for (int q = 0; q <= P.NQ; ++q) {
                    int flag = 0;
                    if (q < P.NQ)
                        for (int i = 0; i < 64; ++i) flag |= P.rhs0[(size_t)q * 64 + i] != 0.0;
                    rowtab[q] = make_int2(P.gptr[q], (q < P.NQ ? P.eneed[q] : ne) | (flag << 30));
                }
                Q.nent = (int)gpack.size();
                Q.tab_bytes = warp_kernel_tab_bytes(P);
                int rc2 = 0;
                rc2 |= upload(h, gpack, &Q.gpack);
                rc2 |= upload(h, rowtab, &Q.rowtab);
                rc2 |= upload(h, P.rhs0, &Q.rhs0);
                rc2 |= upload(h, P.ecoord, &Q.ecoord);
                std::vector<int> elm_fu((size_t)8 * ne);  // band rows of the element dofs, elements in first-use order
                for (int k = 0; k < ne; ++k)
                    for (int a = 0; a < 8; ++a) elm_fu[8 * k + a] = P.elm[8 * P.eord[k] + a];
                rc2 |= upload(h, elm_fu, &Q.elm);
                Q.x_in_smem = (size_t)2 * P.npad <= (size_t)P.R * 36;
                if (rc2) return -2;
                const long long nwarps = (long long)h->num_sms * NW;
                Q.lws_stride = (long long)P.NQ * (kWarpNB + 2) * 64;
                Q.xws_stride = 5LL * P.npad;
                void *pl = nullptr, *px = nullptr;
                CU(cudaMalloc(&pl, (size_t)nwarps * Q.lws_stride * sizeof(double)));
                h->dev_allocs.push_back(pl);
                CU(cudaMalloc(&px, (size_t)nwarps * Q.xws_stride * sizeof(double)));
                h->dev_allocs.push_back(px);
                Q.lws = (double *)pl;
                Q.xws = (double *)px;
                h->warp_nw = NW;
                h->block = NW * 32;
                h->ctas_per_sm = 1;
                h->smem_bytes = smem;
                h->variant = 4;
                h->n_real = n;
                h->PM.NB = kWarpNB;
                h->PM.R = P.R;
                guard.p = nullptr;
                *out = h;
                return 0;
            }
        }
    }

    // ---- on-chip front kernel: band (n x 26 doubles) + five vectors must fit twice per SM
    {
        constexpr int TB = kFrontB, TP = kFrontP, TNT = kFrontNT;
        const FrontPlan plan = plan_front(m, dof2band, n, b, (size_t)prop.sharedMemPerMultiprocessor);
        const bool ok = plan.ok && !force_panel, flip = plan.flip;
        const int pT = plan.pT;
        const int *tip = plan.tip, *er = plan.er;
        auto ori = [&](int g) { return (g < 0) ? g : (flip ? n - 1 - g : g); };
        const size_t fr_smem = front_smem_bytes(n);
        if (ok) {
            const int me = pT + TP, nB = n - me;
            M.b = TB;
            M.ldb = TP;
            M.pT = pT;
            M.nB = nB;
            M.num_sms = prop.multiProcessorCount;
            M.band_in_smem = 1;
            // oriented band row g -> local vector index / band storage
            auto lvi = [&](int g) { return g < me ? g : me + (n - 1 - g); };
            // scratch slot for entries of supported dofs: the last observation slot (vbfem_front_kernel.cuh)
            const unsigned dummy = (unsigned)(((size_t)n * TP + 5 * (size_t)n + 32 + 31) * sizeof(double));
            std::vector<unsigned> eoff((size_t)36 * ne, dummy);
            std::vector<short> ulm((size_t)8 * ne, (short)-1);
            for (int e = 0; e < ne; ++e) {
                int gb[8];
                for (int a = 0; a < 4; ++a)
                    for (int c = 0; c < 2; ++c) gb[2 * a + c] = ori(dof2band[2 * (m->ien[4 * e + a] - 1) + c]);
                for (int a = 0; a < 8; ++a) {
                    if (gb[a] >= 0) ulm[8 * e + a] = (short)lvi(gb[a]);
                    for (int q = 0; q <= a; ++q) {
                        if (gb[a] < 0 || gb[q] < 0) continue;
                        const int lo = std::min(gb[a], gb[q]), hi = std::max(gb[a], gb[q]);
                        // top + middle rows: column lo of the top band; bottom rows: mirrored column of hi
                        const int off = hi < me ? lo * TP + (hi - lo) : me * TP + (n - 1 - hi) * TP + (hi - lo);
                        eoff[36 * e + tri(a, q)] = (unsigned)(off * sizeof(double));
                    }
                }
            }
            std::vector<double> pf_loc(n, 0.0);
            for (int i = 0; i < n; ++i) pf_loc[lvi(ori(dof2band[m->free_dof[i] - 1]))] = m->pf[i];
            for (int k = 0; k < 2; ++k) M.obs_lv[k] = tip[k] >= 0 ? lvi(ori(tip[k])) : -1;
            double ox[4], oy[4];
            for (int a = 0; a < 4; ++a) {
                const int nd = m->ien[4 * (m->obs_ele - 1) + a] - 1;
                ox[a] = m->coord[2 * nd];
                oy[a] = m->coord[2 * nd + 1];
                for (int c = 0; c < 2; ++c) M.obs_lmv[2 * a + c] = er[2 * a + c] >= 0 ? lvi(ori(er[2 * a + c])) : -1;
            }
            for (int q = 0; q < 2; ++q) host_shapef_q4(ox, oy, m->obs_gp[q] - 1, M.obs_nx[q], M.obs_ny[q]);
            int rc2 = 0;
            rc2 |= upload(h, eoff, &M.eoff);
            rc2 |= upload(h, ulm, &M.ulm);
            rc2 |= upload(h, pf_loc, &M.pf_loc);
            {
                std::vector<int> zero(prop.multiProcessorCount, 0);
                const int *p = nullptr;
                rc2 |= upload(h, zero, &p);
                M.sm_ticket = const_cast<int *>(p);
            }
            if (rc2) {
                return -2;
            }
            kernel_fn ks[3] = {fem_front_kernel<TB, TNT, 0>, fem_front_kernel<TB, TNT, 1>,
                               fem_front_kernel<TB, TNT, 2>};
            int nbmin = 1 << 30;
            for (int q = 0; q < 3; ++q) {
                cudaError_t e1 = allow_max_smem(ks[q]);
                int nb = 0;
                if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ks[q], TNT, fr_smem);
                if (e1 != cudaSuccess || nb < 1) {
                    return fail(-3, "front kernel does not fit (%zu bytes of shared memory)", fr_smem);
                }
                nbmin = std::min(nbmin, nb);
                h->kern_front[q] = ks[q];
            }
            h->block = TNT;
            h->ctas_per_sm = std::min(nbmin, 2);
            h->smem_bytes = fr_smem;
            h->variant = 2;
            h->n_real = n;
            guard.p = nullptr;
            *out = h;
            return 0;
        }
    }

    // ---- blocked panel kernel: wide bands (factor streamed to HBM), band order ending at the observed node
    if (getenv("VBFEM_FORCE_GENERIC") == nullptr) {
        PanelPlan P = plan_panel(m, dof2band, n);
        if (P.ok) {
            PanelModel &Q = h->PM;
            Q.n = P.n;
            Q.off = P.off;
            Q.npad = P.npad;
            Q.NQ = P.NQ;
            Q.NB = P.NB;
            Q.R = P.R;
            Q.nub = P.nub;
            Q.obs_loc[0] = P.obs_loc[0];
            Q.obs_loc[1] = P.obs_loc[1];
            Q.o_win = P.o_win;
            Q.o_rhs = P.o_rhs;
            Q.o_lst = P.o_lst;
            Q.o_ke = P.o_ke;
            Q.smem_bytes = P.smem_bytes;
            Q.stages = P.stages;
            Q.o_rec = P.o_rec;
            Q.o_lneg = P.o_lneg;
            Q.rec_stride = P.rec_stride;
            Q.rec_o_src = P.rec_o_src;
            Q.rec_o_dst = P.rec_o_dst;
            for (int w = 0; w <= kPanelNW; ++w) Q.kstart[w] = P.kstart[w];
            for (size_t i = 0; i < P.ub.size(); ++i) Q.ub[i] = P.ub[i];
            const bool dfma = getenv("VBFEM_PANEL_DFMA") != nullptr;  // the DMMA-vs-DFMA comparison
            panel_fn ks[3] = {dfma ? fem_panel_kernel<0, false> : fem_panel_kernel<0, true>,
                              dfma ? fem_panel_kernel<1, false> : fem_panel_kernel<1, true>,
                              dfma ? fem_panel_kernel<2, false> : fem_panel_kernel<2, true>};
            if (P.v2 && getenv("VBFEM_PANEL_V1") == nullptr) {  // second generation: window in registers
                ks[0] = dfma ? fem_panel2_kernel<0, false> : fem_panel2_kernel<0, true>;
                ks[1] = dfma ? fem_panel2_kernel<1, false> : fem_panel2_kernel<1, true>;
                ks[2] = dfma ? fem_panel2_kernel<2, false> : fem_panel2_kernel<2, true>;
                Q.o_wdiag = P.o2_wdiag;
                Q.o_fresh = P.o2_fresh;
                Q.o_vst = P.o2_vst;
                Q.o_big = P.o2_big;
                Q.o_lneg = P.o2_lneg;
                Q.o_rec = P.o2_rec;
                Q.o_lst = P.o2_lst;
                Q.o_ke = P.o2_ke;
                Q.smem_bytes = P.smem2_bytes;
                Q.stages = P.stages2;
                P.smem_bytes = P.smem2_bytes;
            }
            int nbmin = 1 << 30;
            bool fits = true;
            for (int q = 0; q < 3 && fits; ++q) {
                cudaError_t e1 = allow_max_smem(ks[q]);
                int nb = 0;
                if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ks[q], kPanelNT, P.smem_bytes);
                if (e1 != cudaSuccess || nb < 1) {
                    fits = false;
                    cudaGetLastError();
                }
                nbmin = std::min(nbmin, nb);
                h->kern_panel[q] = ks[q];
            }
            if (fits) {
                int rc2 = 0;
                rc2 |= upload(h, P.rec, &Q.rec);
                rc2 |= upload(h, P.eneed, &Q.eneed);
                rc2 |= upload(h, P.ecoord, &Q.ecoord);
                rc2 |= upload(h, P.elm, &Q.elm);
                if (rc2) return -2;
                h->block = kPanelNT;
                h->ctas_per_sm = std::min(nbmin, 2);
                h->smem_bytes = (size_t)P.smem_bytes;
                const long long grid = (long long)h->num_sms * h->ctas_per_sm;
                Q.lws_stride = (long long)P.NQ * (P.NB + 2) * 64;
                Q.xws_stride = 5LL * P.npad;
                void *pl = nullptr, *px = nullptr;
                CU(cudaMalloc(&pl, (size_t)grid * Q.lws_stride * sizeof(double)));
                h->dev_allocs.push_back(pl);
                CU(cudaMalloc(&px, (size_t)grid * Q.xws_stride * sizeof(double)));
                h->dev_allocs.push_back(px);
                Q.lws = (double *)pl;
                Q.xws = (double *)px;
                Q.kews_stride = 36LL * ne;
                void *pk = nullptr;
                CU(cudaMalloc(&pk, (size_t)grid * Q.kews_stride * sizeof(double)));
                h->dev_allocs.push_back(pk);
                Q.kews = (double *)pk;
                h->variant = 3;
                h->n_real = n;
                guard.p = nullptr;
                *out = h;
                return 0;
            }
        }
    }

    // ---- neither fast kernel takes this mesh / observation set-up: the generic kernel serves every mode
    guard.p = nullptr;
    *out = h;
    return 0;
}

extern "C" int vbfem_plan(const vbfem_mesh *m, int64_t smem_per_sm, int64_t *out) {
    if (!m || !out) return fail(-1, "null argument");
    if (m->nnodes <= 0 || m->nele <= 0 || m->nfree <= 0 || !m->coord || !m->ien || !m->free_dof)
        return fail(-1, "incomplete mesh description");
    const int nn = m->nnodes, ndof = 2 * nn;
    for (int i = 0; i < 4 * m->nele; ++i)
        if (m->ien[i] < 1 || m->ien[i] > nn) return fail(-1, "IEN entry %d out of range", m->ien[i]);
    if (m->obs_node < 1 || m->obs_node > nn || m->obs_ele < 1 || m->obs_ele > m->nele)
        return fail(-1, "observation node/element out of range");
    std::vector<char> is_free(ndof, 0);
    for (int i = 0; i < m->nfree; ++i) {
        const int g = m->free_dof[i];
        if (g < 1 || g > ndof) return fail(-1, "free_dof entry %d out of range", g);
        is_free[g - 1] = 1;
    }
    std::vector<int> dof2band;
    const int b = std::max(choose_numbering(m, is_free, dof2band), 1), n = m->nfree;
    if (n > 32767) return fail(-3, "system too large for 16-bit band rows (n = %d)", n);
    const FrontPlan P = plan_front(m, dof2band, n, b, (size_t)(smem_per_sm > 0 ? smem_per_sm : 233472));
    int variant = P.ok ? 2 : 0;
    if (!P.ok && getenv("VBFEM_FORCE_GENERIC") == nullptr && m->pf) {
        const PanelPlan Q = plan_panel(m, dof2band, n);
        if (Q.ok) variant = 3;
    }
    out[1] = n;
    out[2] = b;
    out[7] = 0;
    if (want_warp_kernel() && m->pf) {
        const PanelPlan Q = plan_panel(m, dof2band, n, kWarpNB, warp_kernel_batch());
        // 232448: shared memory a block may opt in to on B200 (sharedMemPerBlockOptin)
        const size_t per_block = (size_t)std::min<int64_t>(smem_per_sm > 0 ? smem_per_sm : 233472, 232448);
        const int NW = warp_kernel_warps(Q, per_block);
        if (warp_kernel_ok(Q, NW, per_block)) {
            const Warp2Plan W2 = warp2_plan(Q, per_block);
            out[0] = 4;
            out[3] = out[4] = out[5] = 0;
            // shared memory of the fused / forward launches (sixteen warps when they fit; Jacobian mode runs twelve)
            out[6] = W2.ok ? (int64_t)(W2.ok16 ? W2.smem16 : W2.smem)
                           : (int64_t)NW * warp_kernel_smem(Q) + warp_kernel_tab_bytes(Q);
            out[7] = W2.ok ? 2 : 1;  // generation of the warp kernel
            return 0;
        }
    }
    out[0] = variant;
    out[3] = P.ok ? P.pT : 0;
    out[4] = P.ok ? n - P.pT - kFrontP : 0;
    out[5] = P.ok && P.flip;
    out[6] = P.ok ? (int64_t)front_smem_bytes(n) : 0;
    return 0;
}

// The panel kernel's host tables for a mesh, GPU-free (tests/panel_emulator.py replays the kernel's
// index logic on them).  which: 0 header (int32: ok, n, off, npad, NQ, NB, R, nub, obs_loc[2], flip,
// entries, smem_bytes, stages, nele, EB), 1 gptr, 2 gdst (u16), 3 gsrc (u16 x 4), 4 eneed, 5 eord,
// 6 rhs0 (f64), 7 elm, 8 ub (u16), 9 kstart.  Returns the byte size of the table (copied when it fits).
extern "C" int64_t vbfem_debug_panel_tables(const vbfem_mesh *m, int which, void *out, int64_t cap_bytes) {
    if (!m) return fail(-1, "null argument");
    const int nn = m->nnodes, ndof = 2 * nn;
    std::vector<char> is_free(ndof, 0);
    for (int i = 0; i < m->nfree; ++i) is_free[m->free_dof[i] - 1] = 1;
    std::vector<int> dof2band;
    choose_numbering(m, is_free, dof2band);
    const bool warp_plan = which >= 100;   // 100 + k: table k of the warp kernel's plan
    if (warp_plan) which -= 100;
    const PanelPlan P = warp_plan ? plan_panel(m, dof2band, m->nfree, kWarpNB, warp_kernel_batch()) : plan_panel(m, dof2band, m->nfree);
    std::vector<int> hdr = {P.ok, P.n, P.off, P.npad, P.NQ, P.NB, P.R, P.nub, P.obs_loc[0], P.obs_loc[1],
                            P.flip, (int)P.gdst.size(), P.smem_bytes, P.stages, P.nele, kPanelEB};
    std::vector<int> ks(P.kstart, P.kstart + kPanelNW + 1);
    const void *src = nullptr;
    int64_t bytes = 0;
    auto pick = [&](const void *p, size_t b) {
        src = p;
        bytes = (int64_t)b;
    };
    switch (which) {
        case 0: pick(hdr.data(), hdr.size() * 4); break;
        case 1: pick(P.gptr.data(), P.gptr.size() * 4); break;
        case 2: pick(P.gdst.data(), P.gdst.size() * 2); break;
        case 3: pick(P.gsrc.data(), P.gsrc.size() * 2); break;
        case 4: pick(P.eneed.data(), P.eneed.size() * 4); break;
        case 5: pick(P.eord.data(), P.eord.size() * 4); break;
        case 6: pick(P.rhs0.data(), P.rhs0.size() * 8); break;
        case 7: pick(P.elm.data(), P.elm.size() * 4); break;
        case 8: pick(P.ub.data(), P.ub.size() * 2); break;
        case 9: pick(ks.data(), ks.size() * 4); break;
        default: return fail(-1, "unknown table %d", which);
    }
    if (out && bytes <= cap_bytes && bytes > 0) memcpy(out, src, (size_t)bytes);
    return bytes;
}

extern "C" void vbfem_destroy(vbfem_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (void *p : h->dev_allocs) cudaFree(p);
    cudaFree(h->jac);
    cudaFree(h->ws_gen);
    cudaFree(h->status);
    cudaFree(h->elbo_f);
    cudaFree(h->elbo_g);
    cudaFree(h->elbo_ysum);
    cudaFree(h->stage);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    for (void *p : h->peer_opened) cudaIpcCloseMemHandle(p);
    cudaFree(h->peer_mail);
    cudaFree(h->peer_state);
    delete h;
}

extern "C" int vbfem_info(const vbfem_t *h, int64_t *out) {
    if (!h || !out) return fail(-1, "null argument");
    for (int i = 0; i < VBFEM_INFO_COUNT; ++i) out[i] = 0;
    out[VBFEM_INFO_NFREE] = h->variant == 2 ? h->n_real : h->M_gen.n;
    out[VBFEM_INFO_HALF_BW] = h->M_gen.b;
    out[VBFEM_INFO_NDOF] = h->M.ndof;
    out[VBFEM_INFO_NELE] = h->M.nele;
    out[VBFEM_INFO_NCOLORS] = h->M.ncolors;
    out[VBFEM_INFO_BAND_IN_SMEM] = h->variant >= 3 ? 0 : h->M.band_in_smem;
    // the warp kernel's forward / fused-adjoint launches (the headline path) may run sixteen warps: report those
    out[VBFEM_INFO_SMEM_BYTES] = h->kern_warp16[0] ? (int64_t)h->warp16_smem : (int64_t)h->smem_bytes;
    out[VBFEM_INFO_CTAS_PER_SM] = h->ctas_per_sm;
    out[VBFEM_INFO_NUM_SMS] = h->num_sms;
    out[VBFEM_INFO_BLOCK_THREADS] = h->kern_warp16[0] ? 512 : h->block;
    out[VBFEM_INFO_KERNEL_VARIANT] = h->variant;
    out[VBFEM_INFO_TWIST_ROW] = h->variant == 2 ? h->M.pT : 0;
    out[VBFEM_INFO_PANEL_BLOCKS] = h->variant >= 3 ? h->PM.NB : 0;
    out[VBFEM_INFO_PANEL_RING] = h->variant >= 3 ? h->PM.R : 0;
    return 0;
}

// Per-sample buffers grow on demand, but never during stream capture (cudaMalloc is illegal there):
// vbfem_reserve sizes them up front.
static int ensure_buf(vbfem_handle *h, void **p, long long *cap, long long need, size_t elem, void *stream,
                      const char *what) {
    if (need <= *cap) return 0;
    cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
    if (stream && cudaStreamIsCapturing((cudaStream_t)stream, &cst) == cudaSuccess && cst != cudaStreamCaptureStatusNone)
        return fail(-6, "%s buffer must grow to %lld entries during stream capture: call vbfem_reserve first", what, need);
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CU(cudaMalloc(p, (size_t)need * elem));
    *cap = need;
    return 0;
}
static int ensure_elbo(vbfem_handle *h, long long nloc, void *stream) {
    if (nloc <= h->elbo_cap) return 0;
    long long c1 = h->elbo_cap, c2 = h->elbo_cap;
    int rc = ensure_buf(h, (void **)&h->elbo_f, &c1, nloc, 2 * sizeof(double), stream, "ELBO");
    if (!rc) rc = ensure_buf(h, (void **)&h->elbo_g, &c2, nloc, 2 * sizeof(double), stream, "ELBO");
    h->elbo_cap = rc ? 0 : nloc;
    return rc;
}
static int ensure_stage(vbfem_handle *h, long long n) {
    if (n <= h->stage_cap) return 0;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    cudaFree(h->stage);
    if (h->pin) cudaFreeHost(h->pin);
    h->stage = h->pin = h->pin_dev = nullptr;
    h->stage_cap = 0;
    CU(cudaMalloc(&h->stage, (size_t)n * 12 * sizeof(double)));
    CU(cudaHostAlloc(&h->pin, (size_t)n * 12 * sizeof(double), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void **)&h->pin_dev, h->pin, 0));
    h->stage_cap = n;
    return 0;
}
static int ensure_ws_gen(vbfem_handle *h, void *stream) {
    if (h->M_gen.band_in_smem || h->ws_gen) return 0;
    long long cap = 0;
    return ensure_buf(h, (void **)&h->ws_gen, &cap, (long long)h->num_sms * h->gen_ctas * h->gen_ws_stride,
                      sizeof(double), stream, "band scratch");
}

extern "C" int vbfem_reserve(vbfem_t *h, int64_t n_samples_max) {
    if (!h || n_samples_max < 0) return fail(-1, "bad argument");
    const long long n = std::max<long long>(n_samples_max, 1);
    int rc = ensure_buf(h, (void **)&h->status, &h->status_cap, n, sizeof(int), nullptr, "status");
    if (!rc) rc = ensure_buf(h, (void **)&h->jac, &h->jac_cap, n, 8 * sizeof(double), nullptr, "Jacobian");
    if (!rc) rc = ensure_elbo(h, n, nullptr);
    if (!rc) rc = ensure_stage(h, n);
    if (!rc && h->variant == 0) rc = ensure_ws_gen(h, nullptr);
    return rc;
}

static int launch(vbfem_handle *h, Args &a, void *stream) {
    if (a.N <= 0) return 0;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_buf(h, (void **)&h->status, &h->status_cap, a.N, sizeof(int), stream, "status");
    if (rc) return rc;
    a.status = h->status;
    h->last_n = a.N;
    const bool fields = (a.mode & kFields) != 0;
    if ((a.mode & kKeep) && !a.ws) {  // Jacobians into the handle's own buffer (vbfem_forward with keep_factor)
        rc = ensure_buf(h, (void **)&h->jac, &h->jac_cap, a.N, 8 * sizeof(double), stream, "Jacobian");
        if (rc) return rc;
        a.ws = h->jac;
    }
    if (a.mode & kKeep) a.ws_stride = 8;
    const int mode = (a.mode & kKeep) ? 2 : ((a.mode & kAdjoint) ? 1 : 0);
    if (h->variant == 2 && !fields) {
        const long long grid = std::min<long long>(a.N, (long long)h->num_sms * h->ctas_per_sm);
#ifdef VBFEM_TIMELINE
        if (!h->timeline) {
            CU(cudaMalloc(&h->timeline, (size_t)h->num_sms * h->ctas_per_sm * 4 * 16 * sizeof(long long)));
        }
        CU(cudaMemsetAsync(h->timeline, 0, (size_t)h->num_sms * h->ctas_per_sm * 4 * 16 * sizeof(long long), st));
        a.timeline = h->timeline;
#endif
        h->kern_front[mode]<<<(unsigned)grid, h->block, h->smem_bytes, st>>>(h->M, a);
    } else if (h->variant == 4 && !fields) {
        const long long grid = std::min<long long>(a.N, (long long)h->num_sms);
#ifdef VBFEM_TIMELINE
        if (!h->timeline) {
            CU(cudaMalloc(&h->timeline, (size_t)h->num_sms * 2 * 4 * 16 * sizeof(long long)));
        }
        CU(cudaMemsetAsync(h->timeline, 0, (size_t)h->num_sms * 4 * 16 * sizeof(long long), st));
        a.timeline = h->timeline;
#endif
        if (h->kern_warp16[mode]) {
            WarpModel wm = h->WM;
            wm.warp_smem = mode == 2 ? h->warp16j_per_warp : h->warp16_per_warp;
            if (mode == 2) wm.win_rows = h->warp16j_rows;
            h->kern_warp16[mode]<<<(unsigned)grid, 512, mode == 2 ? h->warp16j_smem : h->warp16_smem, st>>>(h->M_gen, wm, a);
        } else {
            h->kern_warp[mode]<<<(unsigned)grid, h->block, h->smem_bytes, st>>>(h->M_gen, h->WM, a);
        }
    } else if (h->variant == 3 && !fields) {
        const long long grid = std::min<long long>(a.N, (long long)h->num_sms * h->ctas_per_sm);
#ifdef VBFEM_TIMELINE
        if (!h->timeline) {
            CU(cudaMalloc(&h->timeline, (size_t)h->num_sms * h->ctas_per_sm * 4 * 16 * sizeof(long long)));
        }
        CU(cudaMemsetAsync(h->timeline, 0, (size_t)h->num_sms * h->ctas_per_sm * 4 * 16 * sizeof(long long), st));
        a.timeline = h->timeline;
#endif
        h->kern_panel[mode]<<<(unsigned)grid, h->block, h->smem_bytes, st>>>(h->M_gen, h->PM, a);
    } else {
        const long long grid = std::min<long long>(a.N, (long long)h->num_sms * h->gen_ctas);
        rc = ensure_ws_gen(h, stream);
        if (rc) return rc;
        a.ws_stride = h->gen_ws_stride;
        double *jac = a.ws;
        a.ws = h->ws_gen;
        if (a.mode & kKeep) {
            // Jacobian mode through the generic kernel: one fused forward+adjoint launch per row of
            // d(y0, y1, h0, h1)/dx with a unit cotangent (the rarely used fallback; four factorisations)
            for (int row = 0; row < 4; ++row) {
                Args b = a;
                b.mode = (a.mode & ~kKeep) | kAdjoint;
                b.const_g = 1;
                for (int k = 0; k < 4; ++k) b.gc[k] = (k == row) ? 1.0 : 0.0;
                b.gx = jac;
                b.gx_stride = 8;
                b.gx_off = 2 * row;
                if (row > 0) b.y = b.h = b.f_out = nullptr;
                h->kern<<<(unsigned)grid, h->gen_block, h->gen_smem_bytes, st>>>(h->M_gen, b);
            }
        } else {
            h->kern<<<(unsigned)grid, h->gen_block, h->gen_smem_bytes, st>>>(h->M_gen, a);
        }
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vbfem_forward(vbfem_t *h, int64_t N, const double *x, double *y, double *hh, int keep, void *stream) {
    if (!h || (N > 0 && !x)) return fail(-1, "null argument");
    Args a{};
    a.N = N;
    a.mode = keep ? kKeep : 0;
    a.x = x;
    a.y = y;
    a.h = hh;
    int rc = launch(h, a, stream);
    if (!rc && keep) {
        h->kept_n = N;
        h->kept_ticket = ++h->ticket;
    }
    return rc;
}

extern "C" int64_t vbfem_keep_ticket(const vbfem_t *h) { return h ? h->kept_ticket : 0; }

extern "C" int vbfem_jac_vjp(vbfem_t *h, int64_t N, const double *jac, const double *gy, const double *gh,
                             double *gx, void *stream) {
    if (!h || (N > 0 && (!jac || !gy || !gh || !gx))) return fail(-1, "null argument");
    if (N <= 0) return 0;
    CU(cudaSetDevice(h->device));
    const int nt = 256;
    jac_apply_kernel<<<(unsigned)((N + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(N, jac, 8, gy, gh, gx);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vbfem_backward(vbfem_t *h, int64_t N, const double *gy, const double *gh, double *gx, void *stream) {
    if (!h || (N > 0 && (!gy || !gh || !gx))) return fail(-1, "null argument");
    if (h->kept_n <= 0 || N != h->kept_n)
        return fail(-5, "vbfem_backward: %lld samples requested, the last vbfem_forward(keep_factor=1) kept %lld",
                    (long long)N, (long long)h->kept_n);
    return vbfem_jac_vjp(h, N, h->jac, gy, gh, gx, stream);
}

extern "C" int vbfem_backward_ticket(vbfem_t *h, int64_t ticket, int64_t N, const double *gy, const double *gh,
                                     double *gx, void *stream) {
    if (!h) return fail(-1, "null argument");
    if (ticket <= 0 || ticket != h->kept_ticket)
        return fail(-5, "vbfem_backward: stale ticket %lld (the handle now keeps the Jacobians of ticket %lld)",
                    (long long)ticket, (long long)h->kept_ticket);
    return vbfem_backward(h, N, gy, gh, gx, stream);
}

extern "C" int vbfem_forward_jac(vbfem_t *h, int64_t N, const double *x, double *y, double *hh, double *jac,
                                 void *stream) {
    if (!h || (N > 0 && (!x || !jac))) return fail(-1, "null argument");
    Args a{};
    a.N = N;
    a.mode = kKeep;
    a.x = x;
    a.y = y;
    a.h = hh;
    a.ws = jac;
    return launch(h, a, stream);
}

extern "C" int vbfem_forward_backward(vbfem_t *h, int64_t N, const double *x, const double *gy, const double *gh,
                                      double *y, double *hh, double *gx, void *stream) {
    if (!h || (N > 0 && (!x || !gy || !gh || !gx))) return fail(-1, "null argument");
    Args a{};
    a.N = N;
    a.mode = kAdjoint;
    a.x = x;
    a.y = y;
    a.h = hh;
    a.gy = gy;
    a.gh = gh;
    a.gx = gx;
    return launch(h, a, stream);
}

extern "C" int vbfem_fields(vbfem_t *h, int64_t N, const double *x, const double *emat, double *u, double *sig,
                            double *eps, double *fint, void *stream) {
    if (!h || (N > 0 && !x && !emat)) return fail(-1, "null argument");
    Args a{};
    a.N = N;
    a.mode = kFields;
    a.x = x;
    a.emat = emat;
    a.u_out = u;
    a.sig_out = sig;
    a.eps_out = eps;
    a.fint_out = fint;
    return launch(h, a, stream);
}

extern "C" int vbfem_fields_elementwise(vbfem_t *h, int64_t N, const double *emat, double *y, double *hh, double *u,
                                        double *sig, double *eps, double *fint, void *stream) {
    if (!h || (N > 0 && !emat)) return fail(-1, "null argument");
    Args a{};
    a.N = N;
    a.mode = kFields;
    a.emat = emat;
    a.emat_per_ele = 1;
    a.y = y;
    a.h = hh;
    a.u_out = u;
    a.sig_out = sig;
    a.eps_out = eps;
    a.fint_out = fint;
    return launch(h, a, stream);
}

static int elbo_step1_impl(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end, const double *mu,
                           const double *sig2, const double *e, const double *ybatch, double sig_e, double *sums,
                           double *gmu, double *gsig2, double *f_out, void *stream, bool peer) {
    if (!h || !mu || !sig2 || !e || !ybatch || !sums || (!peer && (!gmu || !gsig2))) return fail(-1, "null argument");
    if (peer && !h->peer_ready) return fail(-6, "no peer mailboxes: vbfem_peer_open / vbfem_peer_connect first");
    if (peer && 3 + 4 * (int64_t)B > h->peer.cap) return fail(-6, "peer mailbox holds %d doubles, the step needs %lld",
                                                              h->peer.cap, 3 + 4 * (long long)B);
    if (B <= 0 || S <= 0 || j_begin < 0 || j_end < j_begin || j_end > (int64_t)B * S)
        return fail(-1, "bad ELBO sample range");
    CU(cudaSetDevice(h->device));
    const long long nloc = j_end - j_begin;
    int rc = ensure_elbo(h, std::max<long long>(nloc, 1), stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ysum_kernel<<<1, 32, 0, st>>>(B, ybatch, h->elbo_ysum);
    Args a{};
    a.N = nloc;
    a.mode = kAdjoint | kElbo;
    a.mu = mu;
    a.sig2 = sig2;
    a.e = e;
    a.ysum = h->elbo_ysum;
    a.B = B;
    a.S = S;
    a.j_begin = j_begin;
    a.gcoef = 1.0 / (sig_e * (double)B * ((double)B * (double)S));
    a.f_out = f_out ? f_out : h->elbo_f;
    a.gx = h->elbo_g;
    rc = launch(h, a, stream);
    if (rc) return rc;
    const int nblk = 1 + B;  // block 0: sums of f; block 1 + b: observation row b
    if (peer)
        elbo_reduce_kernel<true><<<nblk, 256, 0, st>>>(B, S, j_begin, j_end, a.f_out, h->elbo_g, e, sig2, sums, nullptr,
                                                       nullptr, h->peer);
    else
        elbo_reduce_kernel<false><<<nblk, 256, 0, st>>>(B, S, j_begin, j_end, a.f_out, h->elbo_g, e, sig2, sums, gmu,
                                                        gsig2, PeerCtx{});
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vbfem_elbo_step1(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end, const double *mu,
                                const double *sig2, const double *e, const double *ybatch, double sig_e,
                                double *sums, double *gmu, double *gsig2, double *f_out, void *stream) {
    return elbo_step1_impl(h, B, S, j_begin, j_end, mu, sig2, e, ybatch, sig_e, sums, gmu, gsig2, f_out, stream, false);
}

extern "C" int vbfem_elbo_step1_loss(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end, const double *mu,
                                     const double *sig2, const double *log_sig2, const double *e, const double *ybatch,
                                     double sig_e, int32_t allreduce, double *out, void *stream) {
    if (!h || !log_sig2 || !out) return fail(-1, "null argument");
    if (B > 128) return fail(-1, "vbfem_elbo_step1_loss: batch of %d observations, at most 128", B);
    if (!allreduce && (j_begin != 0 || j_end != (int64_t)B * S))
        return fail(-1, "vbfem_elbo_step1_loss without the exchange needs the whole sample range");
    double *tot = out + 1 + 6 * (size_t)B;  // scratch behind the results: the totals [3 + 4B]
    int rc = elbo_step1_impl(h, B, S, j_begin, j_end, mu, sig2, e, ybatch, sig_e, tot, tot + 3, tot + 3 + 2 * (size_t)B,
                             nullptr, stream, allreduce != 0);
    if (rc) return rc;
    elbo_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(B, (double)B * (double)S, sig_e, tot, mu, sig2, log_sig2,
                                                          h->elbo_ysum, out);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vbfem_elbo_step1_allreduce(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                                          const double *mu, const double *sig2, const double *e, const double *ybatch,
                                          double sig_e, double *totals, double *f_out, void *stream) {
    return elbo_step1_impl(h, B, S, j_begin, j_end, mu, sig2, e, ybatch, sig_e, totals, nullptr, nullptr, f_out, stream,
                           true);
}

static int elbo_step2_impl(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end, const double *mu,
                           const double *sig2, const double *e, double *sums, double *h_out, void *stream, bool peer) {
    if (!h || !mu || !sig2 || !e || !sums) return fail(-1, "null argument");
    if (peer && !h->peer_ready) return fail(-6, "no peer mailboxes: vbfem_peer_open / vbfem_peer_connect first");
    if (B <= 0 || S <= 0 || j_begin < 0 || j_end < j_begin || j_end > (int64_t)B * S)
        return fail(-1, "bad ELBO sample range");
    CU(cudaSetDevice(h->device));
    const long long nloc = j_end - j_begin;
    int rc = ensure_elbo(h, std::max<long long>(nloc, 1), stream);
    if (rc) return rc;
    Args a{};
    a.N = nloc;
    a.mode = kElbo;  // forward only: the theta nets are frozen in step 2 (main_custom_training.py:305)
    a.mu = mu;
    a.sig2 = sig2;
    a.e = e;
    a.B = B;
    a.S = S;
    a.j_begin = j_begin;
    a.h = h_out ? h_out : h->elbo_g;
    rc = launch(h, a, stream);
    if (rc) return rc;
    if (peer) hsum_kernel<true><<<1, 256, 0, (cudaStream_t)stream>>>(nloc, a.h, sums, h->peer);
    else hsum_kernel<false><<<1, 256, 0, (cudaStream_t)stream>>>(nloc, a.h, sums, PeerCtx{});
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vbfem_elbo_step2(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end, const double *mu,
                                const double *sig2, const double *e, double *sums, double *h_out, void *stream) {
    return elbo_step2_impl(h, B, S, j_begin, j_end, mu, sig2, e, sums, h_out, stream, false);
}

extern "C" int vbfem_elbo_step2_allreduce(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                                          const double *mu, const double *sig2, const double *e, double *totals,
                                          double *h_out, void *stream) {
    return elbo_step2_impl(h, B, S, j_begin, j_end, mu, sig2, e, totals, h_out, stream, true);
}

// ------------------------------------------------------------------------------------------
// NVLink peer mailboxes (vbfem_peer.cuh): one process per GPU exchanges CUDA IPC handles of its mailbox
// (any host-side all-gather: torch.distributed, MPI, a file); several handles of ONE process pass the
// mailbox addresses instead.
// ------------------------------------------------------------------------------------------
static size_t peer_mail_bytes(int world, int cap) {
    return ((size_t)2 * world * cap) * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
}

extern "C" int vbfem_peer_open(vbfem_t *h, int32_t rank, int32_t world, int32_t cap, void *ipc_handle_out,
                               void **mailbox_out) {
    if (!h) return fail(-1, "null argument");
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || cap < 4)
        return fail(-1, "bad peer geometry (rank %d of %d, %d doubles; at most %d ranks)", rank, world, cap,
                    kPeerMaxWorld);
    if (h->peer_mail) return fail(-6, "peer mailbox already open");
    CU(cudaSetDevice(h->device));
    const size_t bytes = peer_mail_bytes(world, cap);
    CU(cudaMalloc(&h->peer_mail, bytes));
    CU(cudaMemset(h->peer_mail, 0, bytes));
    CU(cudaMalloc(&h->peer_state, 64));
    CU(cudaMemset(h->peer_state, 0, 64));
    CU(cudaDeviceSynchronize());
    h->peer = PeerCtx{};
    h->peer.rank = rank;
    h->peer.world = world;
    h->peer.cap = cap;
    h->peer.seq = (unsigned long long *)h->peer_state;
    h->peer.arrived = (unsigned int *)((char *)h->peer_state + 8);
    h->peer.err = (int *)((char *)h->peer_state + 16);
    const char *to = getenv("VBFEM_PEER_TIMEOUT_MS");
    h->peer.timeout_ns = (unsigned long long)(to ? atoll(to) : 10000) * 1000000ull;
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t ih;
        CU(cudaIpcGetMemHandle(&ih, h->peer_mail));
        memcpy(ipc_handle_out, &ih, sizeof ih);
    }
    if (mailbox_out) *mailbox_out = h->peer_mail;
    return 0;
}

extern "C" int vbfem_peer_connect(vbfem_t *h, const void *ipc_handles, void *const *mailboxes) {
    if (!h || (!ipc_handles && !mailboxes)) return fail(-1, "null argument");
    if (!h->peer_mail) return fail(-6, "vbfem_peer_open first");
    if (h->peer_ready) return fail(-6, "peers already connected");
    CU(cudaSetDevice(h->device));
    for (int r = 0; r < h->peer.world; ++r) {
        if (r == h->peer.rank) {
            h->peer.mail[r] = (double *)h->peer_mail;
        } else if (mailboxes) {  // same process: use the address; enable peer access if it lives on another device
            cudaPointerAttributes at{};
            CU(cudaPointerGetAttributes(&at, mailboxes[r]));
            if (at.type != cudaMemoryTypeDevice) return fail(-1, "mailbox %d is not device memory", r);
            if (at.device != h->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, h->device, at.device));
                if (!can) return fail(-6, "device %d cannot access device %d", h->device, at.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(-2, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
            h->peer.mail[r] = (double *)mailboxes[r];
        } else {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, (const char *)ipc_handles + (size_t)r * sizeof ih, sizeof ih);
            void *p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
            h->peer_opened.push_back(p);
            h->peer.mail[r] = (double *)p;
        }
    }
    h->peer_ready = true;
    return 0;
}

extern "C" int vbfem_peer_allreduce(vbfem_t *h, double *buf, int32_t n, void *stream) {
    if (!h || !buf) return fail(-1, "null argument");
    if (!h->peer_ready) return fail(-6, "no peer mailboxes: vbfem_peer_open / vbfem_peer_connect first");
    if (n < 0 || n > h->peer.cap) return fail(-1, "%d doubles do not fit the mailbox (%d)", n, h->peer.cap);
    CU(cudaSetDevice(h->device));
    peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(h->peer, buf, n);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int64_t vbfem_peer_status(vbfem_t *h) {
    if (!h) return fail(-1, "null argument");
    if (!h->peer_state) return fail(-6, "no peer mailboxes");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    unsigned long long st[3] = {0, 0, 0};
    CU(cudaMemcpy(st, h->peer_state, sizeof st, cudaMemcpyDeviceToHost));
    if ((int)st[2]) return fail(-7, "a peer never arrived at exchange %llu (timed out)", st[0] + 1);
    return (int64_t)st[0];
}

extern "C" int64_t vbfem_status(vbfem_t *h, int32_t *flags_host, int64_t N) {
    if (!h) return fail(-1, "null argument");
    if (N > h->last_n) return fail(-1, "status requested for %lld samples, last launch had %lld", (long long)N,
                                   (long long)h->last_n);
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    std::vector<int> tmp((size_t)std::max<int64_t>(N, 0));
    if (N > 0) CU(cudaMemcpy(tmp.data(), h->status, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (int64_t i = 0; i < N; ++i) {
        if (tmp[i]) ++bad;
        if (flags_host) flags_host[i] = tmp[i];
    }
    return bad;
}

// Host-buffer entry points.  They run on the handle's own stream.  Small batches (the one-sample-at-a-time
// callers: Metropolis steps and KDE evaluators, src/postprocess_lib.py:78-103) skip the staging copies: the
// kernel reads x from, and writes its results to, mapped pinned host memory -- one launch, one synchronise.
constexpr int64_t kMappedMax = 64;

extern "C" int vbfem_forward_host(vbfem_t *h, int64_t N, const double *x, double *y, double *hh) {
    if (!h || (N > 0 && (!x || !y || !hh))) return fail(-1, "null argument");
    if (N <= 0) return 0;
    CU(cudaSetDevice(h->device));
    int rc = ensure_stage(h, N);
    if (rc) return rc;
    cudaStream_t st = h->host_stream;
    memcpy(h->pin, x, (size_t)N * 2 * sizeof(double));
    if (N <= kMappedMax) {
        rc = vbfem_forward(h, N, h->pin_dev, h->pin_dev + 2 * N, h->pin_dev + 4 * N, 0, st);
        if (rc) return rc;
    } else {
        double *dx = h->stage, *dy = dx + 2 * N, *dh = dy + 2 * N;
        CU(cudaMemcpyAsync(dx, h->pin, (size_t)N * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
        rc = vbfem_forward(h, N, dx, dy, dh, 0, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h->pin + 2 * N, dy, (size_t)N * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    memcpy(y, h->pin + 2 * N, (size_t)N * 2 * sizeof(double));
    memcpy(hh, h->pin + 4 * N, (size_t)N * 2 * sizeof(double));
    return 0;
}

extern "C" int vbfem_forward_backward_host(vbfem_t *h, int64_t N, const double *x, const double *gy,
                                           const double *gh, double *y, double *hh, double *gx) {
    if (!h || (N > 0 && (!x || !gy || !gh || !y || !hh || !gx))) return fail(-1, "null argument");
    if (N <= 0) return 0;
    CU(cudaSetDevice(h->device));
    int rc = ensure_stage(h, N);
    if (rc) return rc;
    cudaStream_t st = h->host_stream;
    memcpy(h->pin, x, (size_t)N * 2 * sizeof(double));  // x | gy | gh | y | h | gx
    memcpy(h->pin + 2 * N, gy, (size_t)N * 2 * sizeof(double));
    memcpy(h->pin + 4 * N, gh, (size_t)N * 2 * sizeof(double));
    if (N <= kMappedMax) {
        double *d = h->pin_dev;
        rc = vbfem_forward_backward(h, N, d, d + 2 * N, d + 4 * N, d + 6 * N, d + 8 * N, d + 10 * N, st);
        if (rc) return rc;
    } else {
        double *d = h->stage;
        CU(cudaMemcpyAsync(d, h->pin, (size_t)N * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
        rc = vbfem_forward_backward(h, N, d, d + 2 * N, d + 4 * N, d + 6 * N, d + 8 * N, d + 10 * N, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h->pin + 6 * N, d + 6 * N, (size_t)N * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    memcpy(y, h->pin + 6 * N, (size_t)N * 2 * sizeof(double));
    memcpy(hh, h->pin + 8 * N, (size_t)N * 2 * sizeof(double));
    memcpy(gx, h->pin + 10 * N, (size_t)N * 2 * sizeof(double));
    return 0;
}

extern "C" int vbfem_measure_peaks(int device, double *fp64_tflops, double *copy_gbs) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(-4, "no CUDA device available");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    if (fp64_tflops) {
        const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
        double *out = nullptr;
        CU(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CU(cudaEventRecord(e0));
            dfma_peak_kernel<<<blocks, threads>>>(out, iters);
            CU(cudaEventRecord(e1));
            CU(cudaEventSynchronize(e1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
            if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        cudaFree(out);
        *fp64_tflops = best;
    }
    if (copy_gbs) {
        const size_t bytes = (size_t)1 << 30;
        double4 *a = nullptr, *b = nullptr;
        CU(cudaMalloc(&a, bytes));
        CU(cudaMalloc(&b, bytes));
        CU(cudaMemset(a, 0, bytes));
        double best = 0;
        for (int rep = 0; rep < 6; ++rep) {
            CU(cudaEventRecord(e0));
            copy_kernel<<<prop.multiProcessorCount * 16, 512>>>(a, b, bytes / sizeof(double4));
            CU(cudaEventRecord(e1));
            CU(cudaEventSynchronize(e1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
        }
        cudaFree(a);
        cudaFree(b);
        *copy_gbs = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

#ifdef VBFEM_TIMELINE
// Profiling builds only: clock64 marks of the last sample each CTA processed, [cta][warp][16].
extern "C" int vbfem_debug_timeline(vbfem_t *h, long long *out_host, int64_t max_entries) {
    if (!h || !h->timeline) return fail(-1, "no timeline recorded");
    CU(cudaDeviceSynchronize());
    const int64_t n = std::min<int64_t>(max_entries, (int64_t)h->num_sms * h->ctas_per_sm * 4 * 16);
    CU(cudaMemcpy(out_host, h->timeline, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
    return (int)(h->num_sms * h->ctas_per_sm);
}
#endif
