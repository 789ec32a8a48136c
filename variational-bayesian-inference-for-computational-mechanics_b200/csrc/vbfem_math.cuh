// vbfem_math.cuh -- per-Gauss-point device math of the Cook's-membrane path.
// Each routine cites the upstream statement it reproduces (paths relative to
// nfeng2022/Variational-Bayesian-Inference-for-Computational-Mechanics).
#pragma once
#include <cuda_runtime.h>

namespace vbfem {

// src/fem_preprocess.py:19 -- the 15-digit literal, not 1/sqrt(3).
__device__ constexpr double kSqt13 = 0.577350269189626;
// src/fem_preprocess.py:32-42 -- Pdevs literals.
__device__ constexpr double kTwo3 = 0.666666666666667;
__device__ constexpr double kOne3 = 0.333333333333333;

// Reciprocal without the IEEE slow path: MUFU.RCP64H seed + two Newton steps
// (rel. error ~2^-20 -> 2^-40 -> below one ulp).  Pivots of an SPD band are
// positive normal numbers; non-positive / non-finite pivots are flagged by the
// caller before this result is trusted.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Same seed, cubically convergent correction with a three-deep dependency chain:
// e = 1 - d r0;  r1 = r0 + r0 e (error e^2);  r2 = r1 + r1 e^2 (error e^4 < 2^-80).
__device__ __forceinline__ double fast_rcp3(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double r1 = fma(r, e, r);
    const double e2 = e * e;
    return fma(r1, e2, r1);
}

struct ShapeQ4 {
    double nx[4];  // dN/dx   (shp[0, :])
    double ny[4];  // dN/dy   (shp[1, :])
    double dvol;   // thk * xsj * weight
};

// Q4 shape-function derivatives and Jacobian at Gauss point gp of the 2x2 rule.
// src/fem_preprocess.py:1223-1285 (shapef_tf) with the point table of
// src/fem_preprocess.py:553-558 (order (-,-),(+,-),(+,+),(-,+), weight 1) and
// dvol = thk * jac (src/mat_subroutine_tf.py:76).
__device__ __forceinline__ void shapef_q4(const double (&x)[4], const double (&y)[4], int gp, double thk,
                                          ShapeQ4 &o) {
    const double s0 = (gp == 1 || gp == 2) ? kSqt13 : -kSqt13;
    const double s1 = (gp >= 2) ? kSqt13 : -kSqt13;
    const double sh = 0.5 * s0, th = 0.5 * s1;
    const double sp = 0.5 + sh, tp = 0.5 + th, sm = 0.5 - sh, tm = 0.5 - th;
    const double xo = x[0] - x[1] + x[2] - x[3];
    double xs = -x[0] + x[1] + x[2] - x[3] + xo * s1;
    double xt = -x[0] - x[1] + x[2] + x[3] + xo * s0;
    const double yo = y[0] - y[1] + y[2] - y[3];
    double ys = -y[0] + y[1] + y[2] - y[3] + yo * s1;
    double yt = -y[0] - y[1] + y[2] + y[3] + yo * s0;
    double xsj1 = xs * yt - xt * ys;
    const double xsj = 0.0625 * xsj1;
    xsj1 = (xsj1 != 0.0) ? 1.0 / xsj1 : 1.0;
    xs = (xs + xs) * xsj1;
    xt = (xt + xt) * xsj1;
    ys = (ys + ys) * xsj1;
    yt = (yt + yt) * xsj1;
    const double ytm = yt * tm, ysm = ys * sm, ytp = yt * tp, ysp = ys * sp;
    const double xtm = xt * tm, xsm = xs * sm, xtp = xt * tp, xsp = xs * sp;
    o.nx[0] = -ytm + ysm;
    o.nx[1] = ytm + ysp;
    o.nx[2] = ytp - ysp;
    o.nx[3] = -ytp - ysm;
    o.ny[0] = xtm - xsm;
    o.ny[1] = -xtm - xsp;
    o.ny[2] = -xtp + xsp;
    o.ny[3] = xtp + xsm;
    o.dvol = thk * xsj;  // weight sg2[2] = 1.0
}

// Small-strain measures at a Gauss point (src/mat_subroutine_tf.py:112-145):
// eps_xx, eps_yy and the engineering shear.  ue = (x1,y1,...,x4,y4).
__device__ __forceinline__ void strain_q4(const ShapeQ4 &s, const double (&ue)[8], double &exx, double &eyy,
                                          double &gxy) {
    exx = s.nx[0] * ue[0] + s.nx[1] * ue[2] + s.nx[2] * ue[4] + s.nx[3] * ue[6];
    eyy = s.ny[0] * ue[1] + s.ny[1] * ue[3] + s.ny[2] * ue[5] + s.ny[3] * ue[7];
    gxy = (s.nx[0] * ue[1] + s.nx[1] * ue[3] + s.nx[2] * ue[5] + s.nx[3] * ue[7]) +
          (s.ny[0] * ue[0] + s.ny[1] * ue[2] + s.ny[2] * ue[4] + s.ny[3] * ue[6]);
}

// Consistent tangent on (xx, yy, xy), symmetric 3x3.
struct Tangent {
    double c11, c12, c13, c22, c23, c33;
};

// Material subroutine: isotropic linear elasticity, plane strain
// (src/mat_subroutine_tf.py:333-390): lambda/mu from (E, nu); sig[0:4] =
// Ce @ (exx, eyy, 0, gxy); Ct = Ce restricted to {0,1,3}^2
// (src/mat_subroutine_tf.py:75-76).
struct Lame {
    double lam, mu;
    double szz;  // sigma_zz = szz * (eps_xx + eps_yy): lambda in plane strain, 0 in plane stress
};
__device__ __forceinline__ Lame lame_from_E_nu(double E, double v) {
    Lame m;
    m.lam = v * E / ((1.0 + v) * (1.0 - 2.0 * v));
    m.mu = 0.5 * E / (1.0 + v);
    m.szz = m.lam;
    return m;
}
// Plane stress (src/mat_subroutine.py:283-290): Ce = E / (1 - v^2) [[1, v, 0], [v, 1, 0], [0, 0, (1 - v) / 2]] is
// the plane-strain tangent with lambda replaced by lambda' = v E / (1 - v^2) (lambda' + 2 mu = E / (1 - v^2),
// (1 - v) / 2 * E / (1 - v^2) = mu); sigma_zz = 0.
__device__ __forceinline__ Lame lame_plane_stress(double E, double v) {
    Lame m;
    m.lam = v * E / (1.0 - v * v);
    m.mu = 0.5 * E / (1.0 + v);
    m.szz = 0.0;
    return m;
}
__device__ __forceinline__ void mat_isotropic_plane_strain(const Lame &m, double exx, double eyy, double gxy,
                                                           double (&sig)[4], Tangent &C) {
    const double l2m = m.lam + 2.0 * m.mu;
    sig[0] = l2m * exx + m.lam * eyy;    // + lam*0 + 0*gxy
    sig[1] = m.lam * exx + l2m * eyy;
    sig[2] = m.szz * exx + m.szz * eyy;  // sigma_zz
    sig[3] = m.mu * gxy;
    C.c11 = l2m;
    C.c12 = m.lam;
    C.c13 = 0.0;
    C.c22 = l2m;
    C.c23 = 0.0;
    C.c33 = m.mu;
}
// eps_a^T (dC/dlam) eps_b and eps_a^T (dC/dmu) eps_b for the tangent above
// (used by the adjoint contraction -psi^T (dK/dp) u).
__device__ __forceinline__ void mat_tangent_param_contract(double axx, double ayy, double axy, double bxx,
                                                           double byy, double bxy, double &clam, double &cmu) {
    clam = (axx + ayy) * (bxx + byy);
    cmu = 2.0 * axx * bxx + 2.0 * ayy * byy + axy * bxy;
}

// The reference's von Mises measure (src/fem_postprocess.py:163-185):
// h = sqrt(0.5 * sum((P6 sigma)^2)), P6 = dev3 (+) 0.5 I3 with the truncated
// literals -- deliberately NOT sqrt(3/2 s:s).  Optionally returns dh/dsigma.
__device__ __forceinline__ double von_mises_ref(const double (&sig)[4], double *dsig /* [4] or nullptr */) {
    const double s0 = kTwo3 * sig[0] + (-kOne3) * sig[1] + (-kOne3) * sig[2];
    const double s1 = (-kOne3) * sig[0] + kTwo3 * sig[1] + (-kOne3) * sig[2];
    const double s2 = (-kOne3) * sig[0] + (-kOne3) * sig[1] + kTwo3 * sig[2];
    const double s3 = 0.5 * sig[3];
    const double h = sqrt(0.5 * (s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3));
    if (dsig) {
        const double r = 0.5 / h;
        dsig[0] = r * (kTwo3 * s0 - kOne3 * s1 - kOne3 * s2);
        dsig[1] = r * (-kOne3 * s0 + kTwo3 * s1 - kOne3 * s2);
        dsig[2] = r * (-kOne3 * s0 - kOne3 * s1 + kTwo3 * s2);
        dsig[3] = r * 0.5 * s3;
    }
    return h;
}

// Lower triangle (a >= b) of an 8x8 element matrix, packed.
__device__ __host__ constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }

// Element stiffness kt += dvol * Bm^T Ct Bm (src/mat_subroutine_tf.py:93) with
// the sparse Bm of src/mat_subroutine_tf.py:161-227 multiplied out by hand
// (x-dof column = (nx, 0, ny), y-dof column = (0, ny, nx)).  Lower triangle only.
__device__ __forceinline__ void accumulate_kt(const ShapeQ4 &s, const Tangent &C, double (&ke)[36]) {
    const double c11 = s.dvol * C.c11, c12 = s.dvol * C.c12, c13 = s.dvol * C.c13;
    const double c22 = s.dvol * C.c22, c23 = s.dvol * C.c23, c33 = s.dvol * C.c33;
#pragma unroll
    for (int B = 0; B < 4; ++B) {
        const double nxb = s.nx[B], nyb = s.ny[B];
        // Ct @ Bm[:, 2B]   and   Ct @ Bm[:, 2B+1]
        const double px0 = c11 * nxb + c13 * nyb, px1 = c12 * nxb + c23 * nyb, px2 = c13 * nxb + c33 * nyb;
        const double py0 = c12 * nyb + c13 * nxb, py1 = c22 * nyb + c23 * nxb, py2 = c23 * nyb + c33 * nxb;
#pragma unroll
        for (int A = B; A < 4; ++A) {
            const double nxa = s.nx[A], nya = s.ny[A];
            ke[tri(2 * A, 2 * B)] += nxa * px0 + nya * px2;
            ke[tri(2 * A + 1, 2 * B)] += nya * px1 + nxa * px2;
            ke[tri(2 * A + 1, 2 * B + 1)] += nya * py1 + nxa * py2;
            if (A > B) ke[tri(2 * A, 2 * B + 1)] += nxa * py0 + nya * py2;
        }
    }
}

}  // namespace vbfem
