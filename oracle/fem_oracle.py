"""CPU ORACLE (test infrastructure, NOT product code).

A plain NumPy restatement of the reference's Cook's-membrane hot path, written
from the reference's algorithm and citing the file:line each function follows
(paths are relative to the upstream repository).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl
reference`` legs may import this module; the shipped CUDA path never does.

Parity status: PINNED against the reference's own NumPy twin
(src/fem_solver.py + src/mat_subroutine.py) run unmodified in the build
container -- see tests/golden/make_golden.py and tests/golden/ref_numpy_twin.npz.
The TensorFlow twin (src/fem_solver_tf.py) cannot be imported in this
environment (no tensorflow wheel); it differs from the NumPy twin only in the
linear solver (dense pivoted LU, fem_solver_tf.py:137, vs SuperLU,
fem_solver.py:95), which this oracle follows by using a dense LU solve.

Two forms are provided and cross-checked in tests/test_oracle.py:
  * a loop form (element loop / Gauss loop / dense B matrices) that mirrors the
    reference statement by statement -- slow, used for small cases;
  * a batched torch-float64 form (``TorchOracle``) with the same formulas that
    also yields reverse-mode gradients, standing in for ``tape.gradient``
    (main_custom_training.py:252-256).
"""
from __future__ import annotations

import math

import numpy as np

# fem_preprocess.py:19  (literal, 15 digits -- NOT 1/sqrt(3) to full precision)
SQT13 = 0.577350269189626
# fem_preprocess.py:32-42  Pdevs rows/cols [0,4,8,3,7,2] (fem_postprocess.py:168,179-183)
_TWO3 = 0.666666666666667
_ONE3 = 0.333333333333333
PDEV6 = np.array(
    [
        [_TWO3, -_ONE3, -_ONE3, 0, 0, 0],
        [-_ONE3, _TWO3, -_ONE3, 0, 0, 0],
        [-_ONE3, -_ONE3, _TWO3, 0, 0, 0],
        [0, 0, 0, 0.5, 0, 0],
        [0, 0, 0, 0, 0.5, 0],
        [0, 0, 0, 0, 0, 0.5],
    ],
    dtype=np.float64,
)


# --------------------------------------------------------------------------- mesh
def read_mesh(path):
    """FEAP-style mesh reader.  Follows fem_preprocess.py:114-289
    (header line 2 = nnodes nele nmat ndm ndf nen; sections 'COORdinates ALL',
    'ELEMents ALL', 'BOUNdary conditions', 'FORCe conditions'; the second
    column of every record is dropped, fem_preprocess.py:213-221)."""
    with open(path, "r", newline=None) as f:
        lines = [ln.rstrip("\r\n") for ln in f.read().splitlines()]
    hdr = lines[1].split()
    nnodes, nele, ndm, ndf, nen = int(hdr[0]), int(hdr[1]), int(hdr[3]), int(hdr[4]), int(hdr[5])
    assert ndm == 2 and ndf == 2 and nen == 4

    def find(tag, start=0):
        for i in range(start, len(lines)):
            if lines[i].strip() == tag:
                return i
        return -1

    i = find("COORdinates ALL")
    coord = np.array([[float(t) for t in lines[i + 1 + k].split()] for k in range(nnodes)])
    coord = np.delete(coord[:, :4], 1, axis=1)  # (id, x, y)
    i = find("ELEMents ALL", i)
    elem = np.array([[int(t) for t in lines[i + 1 + k].split()] for k in range(nele)], dtype=np.int64)
    elem = np.delete(elem, 1, axis=1)  # (id, mat, n1..n4)
    conn = elem[:, 2 : 2 + nen]

    def block(tag, start, conv):
        j = find(tag, start)
        rows = []
        if j >= 0:
            k = j + 1
            while k < len(lines) and lines[k].strip():
                rows.append([conv(t) for t in lines[k].split()])
                k += 1
        return j, rows

    j, bc = block("BOUNdary conditions", i, int)
    bc = np.delete(np.array(bc, dtype=np.int64), 1, axis=1) if bc else np.zeros((0, 3), np.int64)
    j2, ld = block("FORCe conditions", max(j, i), float)
    ld = np.delete(np.array(ld, dtype=np.float64), 1, axis=1) if ld else np.zeros((0, 3))
    return dict(nnodes=nnodes, nele=nele, coord=coord, conn=conn, bc=bc, load=ld)


def assign_dof(mesh):
    """DOF maps.  Follows fem_preprocess.py:291-443: ID[c, n] = 2 n + c + 1
    (1-based), LM[:, e] = ID[:, IEN[e]].flatten('F'), supports from the
    boundary flags, Pf = nodal loads gathered at the free dofs (x loads only
    where the x entry is non-zero, same for y, fem_preprocess.py:362-371)."""
    nnodes, nele = mesh["nnodes"], mesh["nele"]
    ndof = 2 * nnodes
    ID = np.arange(1, ndof + 1).reshape(nnodes, 2).T
    conn = mesh["conn"]
    supp = []
    for row in mesh["bc"]:
        n = int(row[0])
        if row[1] == 1:
            supp.append(ID[0, n - 1])
        if row[2] == 1:
            supp.append(ID[1, n - 1])
    supp_dof = np.unique(np.array(supp, dtype=np.int64))
    free_dof = np.setdiff1d(np.arange(1, ndof + 1), supp_dof)
    LM = np.zeros((8, nele), dtype=np.int64)
    for e in range(nele):
        LM[:, e] = ID[:, conn[e] - 1].flatten(order="F")
    P = np.zeros(ndof)
    for row in mesh["load"]:
        n = int(row[0])
        if row[1] != 0:
            P[ID[0, n - 1] - 1] += row[1]
        if row[2] != 0:
            P[ID[1, n - 1] - 1] += row[2]
    return dict(ID=ID, IEN=conn.copy(), LM=LM, free_dof=free_dof, supp_dof=supp_dof, ndof=ndof,
                Pf=P[free_dof - 1].copy())


def cook_mesh_text(nx, ny, total_load=50.0, digits=6):
    """Cook's membrane nx x ny written in the reference's FEAP text format
    (same bilinear map that reproduces Armero_cooksm_20x10.txt; SURVEY 8d
    config 4): corners (0,0),(48,44),(48,60),(0,44), ids row-major nx+1 per
    row, clamp column i=0, load in +y lumped on the right edge.  With the
    shipped file's 7 significant digits (digits=6) the 20x10 coordinates parse
    to the same doubles as the reference mesh."""
    out = ["FEAP * * PLANE strain problem", f"{(nx+1)*(ny+1):10d}{nx*ny:10d}{1:10d}{2:10d}{2:10d}{4:10d}", " ", "",
           "COORdinates ALL"]
    for j in range(ny + 1):
        for i in range(nx + 1):
            xi, eta = i / nx, j / ny
            X = 48.0 * xi
            Y = 44.0 * xi + eta * (44.0 * (1.0 - xi) + 16.0 * xi)
            out.append(f"{j*(nx+1)+i+1:9d} 0 {X: .{digits}E} {Y: .{digits}E}")
    out += ["", "ELEMents ALL"]
    e = 0
    for j in range(ny):
        for i in range(nx):
            e += 1
            n1 = j * (nx + 1) + i + 1
            out.append(f"{e:6d}   0     1 {n1:6d} {n1+1:6d} {n1+nx+2:6d} {n1+nx+1:6d}")
    out += ["", "BOUNdary conditions"]
    for j in range(ny + 1):
        out.append(f"{j*(nx+1)+1:10d}   0   1   1")
    out += ["", "FORCe conditions"]
    for j in range(ny + 1):
        w = total_load / ny * (0.5 if j in (0, ny) else 1.0)
        out.append(f"{j*(nx+1)+nx+1:9d} 0 {0.0: .{digits}E} {w: .{digits}E}")
    out += [" ", "END", ""]
    return "\n".join(out)


# ----------------------------------------------------------------------- element
def gauss_2x2():
    """fem_preprocess.py:553-558 (int2d, l=2): points g*(lr,lz), weights 1."""
    lr = np.array([-1, 1, 1, -1], dtype=np.float64)
    lz = np.array([-1, -1, 1, 1], dtype=np.float64)
    sg = np.zeros((3, 4))
    sg[0], sg[1], sg[2] = SQT13 * lr, SQT13 * lz, 1.0
    return sg


def shapef(s, xl):
    """fem_preprocess.py:904-971 (= shapef_tf 1223-1285), flg = 0."""
    sh, th = 0.5 * s[0], 0.5 * s[1]
    sp, tp, sm, tm = 0.5 + sh, 0.5 + th, 0.5 - sh, 0.5 - th
    xo = xl[0, 0] - xl[0, 1] + xl[0, 2] - xl[0, 3]
    xs = -xl[0, 0] + xl[0, 1] + xl[0, 2] - xl[0, 3] + xo * s[1]
    xt = -xl[0, 0] - xl[0, 1] + xl[0, 2] + xl[0, 3] + xo * s[0]
    yo = xl[1, 0] - xl[1, 1] + xl[1, 2] - xl[1, 3]
    ys = -xl[1, 0] + xl[1, 1] + xl[1, 2] - xl[1, 3] + yo * s[1]
    yt = -xl[1, 0] - xl[1, 1] + xl[1, 2] + xl[1, 3] + yo * s[0]
    xsj1 = xs * yt - xt * ys
    xsj = 0.0625 * xsj1
    xsj1 = 1.0 / xsj1 if xsj1 != 0.0 else 1.0
    xs, xt, ys, yt = (xs + xs) * xsj1, (xt + xt) * xsj1, (ys + ys) * xsj1, (yt + yt) * xsj1
    ytm, ysm, ytp, ysp = yt * tm, ys * sm, yt * tp, ys * sp
    xtm, xsm, xtp, xsp = xt * tm, xs * sm, xt * tp, xs * sp
    shp = np.zeros((3, 4))
    shp[0] = [-ytm + ysm, ytm + ysp, ytp - ysp, -ytp - ysm]
    shp[1] = [xtm - xsm, -xtm - xsp, -xtp + xsp, xtp + xsm]
    shp[2] = [sm * tm, sp * tm, sp * tp, sm * tp]
    return shp, xsj


def isotropic_elasticity(eps, E, v, stype=2):
    """mat_subroutine_tf.py:333-390 (the plane-strain branch, which the TF file executes unconditionally):
    lambda/mu, 4x4 Ce, sig[0:4] = Ce @ eps[0:4], Ct[{0,1,3}^2] = Ce[{0,1,3}^2]; with stype = 1 the NumPy
    twin's plane-stress branch (mat_subroutine.py:283-290).  Returns sig, Ct, eps33."""
    sig = np.zeros(6)
    Ct = np.zeros((6, 6))
    idx = [0, 1, 3]
    eps33 = None
    if stype == 1:
        Ce = E / (1 - v ** 2) * np.array([[1, v, 0], [v, 1, 0], [0, 0, (1 - v) / 2]])
        sig[idx] = Ce @ eps[idx]
        eps33 = -v / (1 - v) * (eps[0] + eps[1])
        Ct[np.ix_(idx, idx)] = Ce
    else:
        l = v * E / ((1 + v) * (1 - 2 * v))
        mu = 0.5 * E / (1 + v)
        Ce = np.array([[l + 2 * mu, l, l, 0], [l, l + 2 * mu, l, 0], [l, l, l + 2 * mu, 0], [0, 0, 0, mu]])
        sig[0:4] = Ce @ eps[0:4]
        Ct[np.ix_(idx, idx)] = Ce[np.ix_(idx, idx)]
    return sig, Ct, eps33


def solid_2d(ul, xl, E, v, thk, stype=2, body=(0.0, 0.0)):
    """mat_subroutine_tf.py:23-110: 2x2 Gauss loop, strain (112-145), plane
    strain eps[2]=0 (54-56), material, Ct -> [0,1,3]^2 (75-76), dvol=thk*jac,
    p += dvol*Bm^T sig[0,1,3] - dvol*Nm^T body (147-159), kt += dvol*Bm^T Ct Bm,
    Bm rows (dNx | dNy | dNy,dNx), Nm rows (N at x dofs | N at y dofs) (161-227).  stype = 1: plane stress,
    eps[2] = eps33 after the material call (mat_subroutine.py:51-52)."""
    sg = gauss_2x2()
    p = np.zeros(8)
    kt = np.zeros((8, 8))
    eps_out = np.zeros((6, 4))
    sig_out = np.zeros((6, 4))
    body = np.asarray(body, dtype=np.float64)
    for ipt in range(4):
        shp, xsj = shapef(sg[0:2, ipt], xl)
        jac = xsj * sg[2, ipt]
        eps = np.zeros(6)
        eps[0] = shp[0] @ ul[0]
        eps[1] = shp[1] @ ul[1]
        eps[3] = shp[0] @ ul[1] + shp[1] @ ul[0]
        eps[2] = 0.0
        sig, Ct, eps33 = isotropic_elasticity(eps, E, v, stype)
        if stype == 1:
            eps[2] = eps33
        Ct3 = Ct[np.ix_([0, 1, 3], [0, 1, 3])]
        dvol = thk * jac
        Bm = np.zeros((3, 8))
        Nm = np.zeros((2, 8))
        for i in range(4):
            Bm[0, 2 * i] = shp[0, i]
            Bm[1, 2 * i + 1] = shp[1, i]
            Bm[2, 2 * i] = shp[1, i]
            Bm[2, 2 * i + 1] = shp[0, i]
            Nm[0, 2 * i] = shp[2, i]
            Nm[1, 2 * i + 1] = shp[2, i]
        p += dvol * (Bm.T @ sig[[0, 1, 3]]) - dvol * (Nm.T @ body)
        kt += dvol * (Bm.T @ Ct3 @ Bm)
        eps_out[:, ipt] = eps
        sig_out[:, ipt] = sig
    return p, kt, eps_out, sig_out


# ------------------------------------------------------------------------ solver
class LoopOracle:
    """Statement-by-statement restatement of fem_solver_tf.py:86-185,229-341."""

    def __init__(self, mesh, dof, thk=10.0, stype=2, body=(0.0, 0.0)):
        self.mesh, self.dof, self.thk = mesh, dof, thk
        self.stype, self.body = stype, body
        self.xy = mesh["coord"][:, 1:3]

    def assemble(self, u, E, v):
        """fem_solver_tf.py:229-341: element loop, gather ul by LM, xl by IEN,
        scatter (duplicates add) into dense Kg and F_int."""
        d, m = self.dof, self.mesh
        ndof, nele = d["ndof"], m["nele"]
        Kg = np.zeros((ndof, ndof))
        Fint = np.zeros(ndof)
        strain = np.zeros((6, 4, nele))
        stress = np.zeros((6, 4, nele))
        for e in range(nele):
            lm = d["LM"][:, e] - 1
            ul = u[lm].reshape(4, 2).T
            xl = self.xy[d["IEN"][e] - 1].T
            Ee, ve = (E[e], v[e]) if np.ndim(E) else (E, v)   # one material, or one per element
            p, kt, eps, sig = solid_2d(ul, xl, Ee, ve, self.thk, self.stype, self.body)
            Fint[lm] += p
            Kg[np.ix_(lm, lm)] += kt
            strain[:, :, e], stress[:, :, e] = eps, sig
        return Kg, Fint, strain, stress

    def solve(self, E, v):
        """fem_solver_tf.py:86-185: zero predictor, assemble, one dense solve on
        the free block, update, assemble again for stresses; exit_flag = 1."""
        d = self.dof
        free = d["free_dof"] - 1
        u = np.zeros(d["ndof"])
        Kg, Fint, _, _ = self.assemble(u, E, v)
        Fext = np.zeros(d["ndof"])
        Fext[free] = d["Pf"]
        R = Fint[free] - Fext[free]
        duf = np.linalg.solve(Kg[np.ix_(free, free)], -R)
        u[free] = u[free] + duf
        _, Fint, strain, stress = self.assemble(u, E, v)
        return u, Fint, strain, stress


def von_mises(stress_e, nipt_id):
    """fem_postprocess.py:163-185: sqrt(0.5*sum((Pdev6 @ sigma)^2)) at the
    requested Gauss points (1-based) -- the reference's non-standard formula."""
    s = stress_e[:, np.asarray(nipt_id) - 1]
    return np.sqrt(0.5 * np.sum((PDEV6 @ s) ** 2, axis=0))


def theta_to_material(x, theta_mean, theta_std):
    """data_generation_2sam_more_loss.py:181-186."""
    E = np.exp(theta_std[0] * x[..., 0] + theta_mean[0])
    v = 0.5 / (1.0 + np.exp(-theta_std[1] * x[..., 1] - theta_mean[1]))
    return E, v


def fem_fh_loop(oracle, x, theta_mean, theta_std, node_id=231, ele_id=12, nipt_id=(1, 3)):
    """data_generation_2sam_more_loss.py:169-192 for a batch x[N,2] (serial)."""
    x = np.atleast_2d(x)
    y = np.zeros((x.shape[0], 2))
    h = np.zeros((x.shape[0], 2))
    for i in range(x.shape[0]):
        E, v = theta_to_material(x[i], theta_mean, theta_std)
        u, _, _, stress = oracle.solve(float(E), float(v))
        y[i] = u[2 * node_id - 2 : 2 * node_id]
        h[i] = von_mises(stress[:, :, ele_id - 1], nipt_id)
    return y, h


# ------------------------------------------------------------ batched torch form
class TorchOracle:
    """Batched float64 torch-CPU form of the same formulas (dense Kg, dense LU
    solve) whose autograd stands in for TensorFlow's tape.gradient
    (main_custom_training.py:252-256).  Geometry (shape-function derivatives,
    dvol) is evaluated once with ``shapef`` above."""

    def __init__(self, mesh, dof, thk=10.0, theta_mean=(math.log(20.0), 0.0), theta_std=(0.1, 0.015),
                 node_id=231, ele_id=12, nipt_id=(1, 3)):
        import torch

        self.torch = torch
        self.mesh, self.dof = mesh, dof
        nele = mesh["nele"]
        xy = mesh["coord"][:, 1:3]
        sg = gauss_2x2()
        B = np.zeros((nele, 4, 3, 8))
        dvol = np.zeros((nele, 4))
        for e in range(nele):
            xl = xy[dof["IEN"][e] - 1].T
            for g in range(4):
                shp, xsj = shapef(sg[0:2, g], xl)
                dvol[e, g] = thk * xsj * sg[2, g]
                for i in range(4):
                    B[e, g, 0, 2 * i] = shp[0, i]
                    B[e, g, 1, 2 * i + 1] = shp[1, i]
                    B[e, g, 2, 2 * i] = shp[1, i]
                    B[e, g, 2, 2 * i + 1] = shp[0, i]
        self.B = torch.from_numpy(B)
        self.dvol = torch.from_numpy(dvol)
        self.lm = torch.from_numpy(dof["LM"].T.copy() - 1)  # [nele, 8]
        self.free = torch.from_numpy(dof["free_dof"] - 1)
        self.Pf = torch.from_numpy(dof["Pf"].copy())
        self.ndof = dof["ndof"]
        self.theta_mean = torch.tensor(theta_mean, dtype=torch.float64)
        self.theta_std = torch.tensor(theta_std, dtype=torch.float64)
        self.node_id, self.ele_id = node_id, ele_id
        self.nipt = torch.tensor([g - 1 for g in nipt_id])
        self.P6 = torch.from_numpy(PDEV6)

    def material(self, x):
        t = self.torch
        E = t.exp(self.theta_std[0] * x[:, 0] + self.theta_mean[0])
        v = 0.5 / (1.0 + t.exp(-self.theta_std[1] * x[:, 1] - self.theta_mean[1]))
        return E, v

    def fields(self, x):
        """x[N,2] -> u[N,ndof], strain/stress[N,6,4,nele] (all differentiable)."""
        t = self.torch
        N = x.shape[0]
        E, v = self.material(x)
        lam = v * E / ((1 + v) * (1 - 2 * v))
        mu = 0.5 * E / (1 + v)
        z = t.zeros_like(lam)
        C3 = t.stack([t.stack([lam + 2 * mu, lam, z], -1), t.stack([lam, lam + 2 * mu, z], -1),
                      t.stack([z, z, mu], -1)], -2)  # [N,3,3]
        # kt[n,e] = sum_g dvol * B^T C B
        CB = t.einsum("nij,egjk->negik", C3, self.B)
        Ke = t.einsum("eg,egia,negib->neab", self.dvol, self.B, CB)
        nele = self.B.shape[0]
        rows = self.lm[:, :, None].expand(nele, 8, 8).reshape(-1)
        cols = self.lm[:, None, :].expand(nele, 8, 8).reshape(-1)
        Kg = t.zeros((N, self.ndof * self.ndof), dtype=t.float64)
        Kg = Kg.index_add(1, rows * self.ndof + cols, Ke.reshape(N, -1)).reshape(N, self.ndof, self.ndof)
        Kff = Kg[:, self.free][:, :, self.free]
        uf = t.linalg.solve(Kff, self.Pf.expand(N, -1).unsqueeze(-1)).squeeze(-1)
        u = t.zeros((N, self.ndof), dtype=t.float64).index_copy(1, self.free, uf)
        ue = u[:, self.lm]  # [N,nele,8]
        eps3 = t.einsum("egia,nea->negi", self.B, ue)  # [N,nele,4,3] (xx,yy,gxy)
        exx, eyy, gxy = eps3[..., 0], eps3[..., 1], eps3[..., 2]
        l_, m_ = lam[:, None, None], mu[:, None, None]
        sxx = (l_ + 2 * m_) * exx + l_ * eyy
        syy = l_ * exx + (l_ + 2 * m_) * eyy
        szz = l_ * exx + l_ * eyy
        sxy = m_ * gxy
        zz = t.zeros_like(sxx)
        stress = t.stack([sxx, syy, szz, sxy, zz, zz], 1).permute(0, 1, 3, 2)  # [N,6,4,nele]
        strain = t.stack([exx, eyy, zz, gxy, zz, zz], 1).permute(0, 1, 3, 2)
        return u, strain, stress

    def fem_fh(self, x):
        """Batched MeasurementData.fem_fh_fun_loop_rev: x[N,2] -> y[N,2], h[N,2]."""
        t = self.torch
        u, _, stress = self.fields(x)
        y = u[:, 2 * self.node_id - 2 : 2 * self.node_id]
        s = stress[:, :, :, self.ele_id - 1][:, :, self.nipt]  # [N,6,2]
        h = t.sqrt(0.5 * t.sum(t.einsum("ij,njk->nik", self.P6, s) ** 2, dim=1))
        return y, h

    def fem_fh_chunked(self, x, chunk=256):
        """``fem_fh`` for large batches: the dense per-sample matrices of a chunk (0.9 GB per 512 samples,
        several times that with the autograd tape) exist only while the chunk is processed; backward
        recomputes the chunk.  Same numbers and gradients as ``fem_fh``."""
        t = self.torch
        oracle = self

        class Chunked(t.autograd.Function):
            @staticmethod
            def forward(ctx, xx):
                ctx.save_for_backward(xx)
                ys, hs = [], []
                with t.no_grad():
                    for i in range(0, xx.shape[0], chunk):
                        y, h = oracle.fem_fh(xx[i:i + chunk])
                        ys.append(y)
                        hs.append(h)
                return t.cat(ys), t.cat(hs)

            @staticmethod
            def backward(ctx, gy, gh):
                (xx,) = ctx.saved_tensors
                out = []
                for i in range(0, xx.shape[0], chunk):
                    with t.enable_grad():
                        xc = xx[i:i + chunk].detach().requires_grad_(True)
                        y, h = oracle.fem_fh(xc)
                        (g,) = t.autograd.grad((y * gy[i:i + chunk]).sum() + (h * gh[i:i + chunk]).sum(), xc)
                    out.append(g)
                return t.cat(out)

        return Chunked.apply(x)

    def vjp(self, x_np, gy_np, gh_np):
        t = self.torch
        x = t.tensor(x_np, dtype=t.float64, requires_grad=True)
        y, h = self.fem_fh(x)
        (gx,) = t.autograd.grad((y * t.tensor(gy_np)).sum() + (h * t.tensor(gh_np)).sum(), x)
        return y.detach().numpy(), h.detach().numpy(), gx.numpy()


# ----------------------------------------------------------------- ELBO (step 1)
def elbo_step1_torch(torch_oracle, y_batch, mu, sig2, e_data, sig_e):
    """main_custom_training.py:183-235 with log_theta_sig = log(sig2):
    loss = term1 - term2 - term3, including the [B, B*S] broadcast of
    (y_point - f_data) at main_custom_training.py:205,210-214.
    Returns (loss, term1, term2, term3) as torch scalars."""
    t = torch_oracle.torch
    d = mu.shape[-1]
    dy = y_batch.shape[-1]
    term1 = -0.5 * t.mean(t.sum(t.log(sig2), dim=-1), dim=0) - 0.5 * d * math.log(2.0 * math.pi) - 0.5 * d
    std = t.sqrt(sig2).unsqueeze(1)
    theta = (e_data * std + mu.unsqueeze(1)).reshape(-1, d)
    f, _ = torch_oracle.fem_fh(theta) if theta.shape[0] <= 512 else torch_oracle.fem_fh_chunked(theta)  # [B*S, 2]
    l1 = -0.5 * dy * math.log(2.0 * math.pi * sig_e)
    l2 = -0.5 / sig_e * t.sum((y_batch.unsqueeze(1) - f) ** 2, dim=-1)  # [B, B*S]
    term2 = l1 + t.mean(l2)
    term3 = -0.5 * d * math.log(2.0 * math.pi) - 0.5 * t.mean(t.sum(sig2 + mu ** 2, dim=-1), dim=0)
    return term1 - term2 - term3, term1, term2, term3


def elbo_step2_torch(torch_oracle, mu, sig2, z_mean, z_sig, log_z_sig, logz_mean_post, logz_sig_post, e_data,
                     sig_eta, alpha=1.0):
    """main_custom_training.py:338-384 statement by statement: term4, term5 (with the [B, B*S]
    broadcast of h_data against z_mean_point[B, 1, .] at main_custom_training.py:347-364), add_loss;
    loss = (term4 - term5) * alpha + add_loss.  The theta nets are frozen (main_custom_training.py:305):
    mu, sig2 carry no gradient."""
    t = torch_oracle.torch
    zd = z_mean.shape[-1]
    term4 = t.mean(-0.5 * t.sum(log_z_sig, dim=-1) - t.sum(z_mean, dim=-1)) - 0.5 * zd * math.log(2.0 * math.pi) \
        - 0.5 * zd
    std = t.sqrt(sig2).unsqueeze(1)
    theta = (e_data * std + mu.unsqueeze(1)).reshape(-1, mu.shape[-1])
    _, h = torch_oracle.fem_fh(theta.detach())           # [B*S, 2]
    zm, zs = z_mean.unsqueeze(1), z_sig.unsqueeze(1)      # [B, 1, 2]
    l1 = -0.5 / sig_eta * t.sum(t.exp(2.0 * zm + 2.0 * zs), dim=-1)                       # [B, 1]
    l2 = -0.5 / sig_eta * t.sum(-2.0 * (h * t.exp(zm + 0.5 * zs)) + h ** 2, dim=-1)       # [B, B*S]
    l3 = -0.5 * zd * math.log(2.0 * math.pi * sig_eta)
    term5 = t.mean(l1 + l2) + l3
    add = t.mean((z_mean - logz_mean_post) ** 2) + t.mean((z_sig - logz_sig_post) ** 2)
    return (term4 - term5) * alpha + add, term4, term5, add


# ------------------------------------------------------ sparse form (large meshes)
class SparseOracle:
    """The reference NumPy twin's own solver route -- CSR assembly by
    (loc_i, loc_j) triplets and ``scipy.sparse.linalg.spsolve``
    (fem_solver.py:68-126, 244-250) -- vectorised over elements, for meshes
    where the dense LU of the TF path (6560^2 at 80x40) is too slow on a CPU.
    Same element formulas as ``solid_2d``."""

    def __init__(self, mesh, dof, thk=10.0, theta_mean=(math.log(20.0), 0.0), theta_std=(0.1, 0.015)):
        self.mesh, self.dof, self.thk = mesh, dof, thk
        self.theta_mean, self.theta_std = np.asarray(theta_mean), np.asarray(theta_std)
        nele = mesh["nele"]
        xy = mesh["coord"][:, 1:3]
        sg = gauss_2x2()
        self.B = np.zeros((nele, 4, 3, 8))
        self.dvol = np.zeros((nele, 4))
        for e in range(nele):
            xl = xy[dof["IEN"][e] - 1].T
            for g in range(4):
                shp, xsj = shapef(sg[0:2, g], xl)
                self.dvol[e, g] = thk * xsj * sg[2, g]
                self.B[e, g, 0, 0::2] = shp[0]
                self.B[e, g, 1, 1::2] = shp[1]
                self.B[e, g, 2, 0::2] = shp[1]
                self.B[e, g, 2, 1::2] = shp[0]
        self.lm = dof["LM"].T - 1

    def solve(self, E, v):
        import scipy.sparse as sp
        from scipy.sparse.linalg import spsolve

        lam = v * E / ((1 + v) * (1 - 2 * v))
        mu = 0.5 * E / (1 + v)
        C3 = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
        Ke = np.einsum("eg,egia,ij,egjb->eab", self.dvol, self.B, C3, self.B)
        ndof = self.dof["ndof"]
        rows = np.repeat(self.lm[:, :, None], 8, axis=2).ravel()
        cols = np.repeat(self.lm[:, None, :], 8, axis=1).ravel()
        K = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(ndof, ndof))
        free = self.dof["free_dof"] - 1
        uf = spsolve(K[free][:, free].tocsc(), self.dof["Pf"])
        u = np.zeros(ndof)
        u[free] = uf
        eps3 = np.einsum("egia,ea->egi", self.B, u[self.lm])
        exx, eyy, gxy = eps3[..., 0], eps3[..., 1], eps3[..., 2]
        z = np.zeros_like(exx)
        stress = np.stack([(lam + 2 * mu) * exx + lam * eyy, lam * exx + (lam + 2 * mu) * eyy,
                           lam * (exx + eyy), mu * gxy, z, z]).transpose(0, 2, 1)  # [6,4,nele]
        return u, stress

    def fem_fh(self, x, node_id, ele_id, nipt_id=(1, 3)):
        x = np.atleast_2d(x)
        y, h = np.zeros((len(x), 2)), np.zeros((len(x), 2))
        for i, xi in enumerate(x):
            E, v = theta_to_material(xi, self.theta_mean, self.theta_std)
            u, stress = self.solve(float(E), float(v))
            y[i] = u[2 * node_id - 2: 2 * node_id]
            h[i] = von_mises(stress[:, :, ele_id - 1], nipt_id)
        return y, h

    def vjp(self, x, gy, gh, node_id, ele_id, nipt_id=(1, 3)):
        """x[N,2], gy[N,2], gh[N,2] -> y, h, gx = d(sum gy*y + sum gh*h)/dx: what tape.gradient
        (main_custom_training.py:252-256) returns, derived per sample as the discrete adjoint of the
        sparse solve (SuperLU factor reused for K psi = w) with torch autograd for the local pieces
        (theta -> E, nu -> lambda, mu; strain -> stress -> von Mises measure at the observed points).
        Checked against TorchOracle.vjp (dense LU + full autograd) in tests/test_oracle.py."""
        import scipy.sparse as sp
        import torch as t
        from scipy.sparse.linalg import splu

        x, gy, gh = np.atleast_2d(x), np.atleast_2d(gy), np.atleast_2d(gh)
        ndof = self.dof["ndof"]
        free = self.dof["free_dof"] - 1
        rows = np.repeat(self.lm[:, :, None], 8, axis=2).ravel()
        cols = np.repeat(self.lm[:, None, :], 8, axis=1).ravel()
        Cl = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]])   # dC3/dlambda
        Cm = np.array([[2.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 1.0]])   # dC3/dmu
        if not hasattr(self, "_Kel"):
            self._Kel = np.einsum("eg,egia,ij,egjb->eab", self.dvol, self.B, Cl, self.B)
            self._Kem = np.einsum("eg,egia,ij,egjb->eab", self.dvol, self.B, Cm, self.B)
        Bo = t.tensor(self.B[ele_id - 1][[g - 1 for g in nipt_id]])           # [2,3,8]
        lmo = self.lm[ele_id - 1]
        P6 = t.tensor(PDEV6)
        y, h, gx = np.zeros((len(x), 2)), np.zeros((len(x), 2)), np.zeros((len(x), 2))
        for i in range(len(x)):
            xt = t.tensor(x[i], dtype=t.float64, requires_grad=True)
            E = t.exp(self.theta_std[0] * xt[0] + self.theta_mean[0])
            v = 0.5 / (1.0 + t.exp(-self.theta_std[1] * xt[1] - self.theta_mean[1]))
            lam = v * E / ((1 + v) * (1 - 2 * v))
            mu = 0.5 * E / (1 + v)
            lam_f, mu_f = float(lam.detach()), float(mu.detach())
            K = sp.csr_matrix(((lam_f * self._Kel + mu_f * self._Kem).ravel(), (rows, cols)), shape=(ndof, ndof))
            lu = splu(K[free][:, free].tocsc())
            u = np.zeros(ndof)
            u[free] = lu.solve(self.dof["Pf"])
            # local observation function of (u_e, lambda, mu)
            ue = t.tensor(u[lmo], requires_grad=True)
            eps = t.einsum("gia,a->gi", Bo, ue)                                   # [2,3] (xx, yy, gxy)
            exx, eyy, gxy = eps[:, 0], eps[:, 1], eps[:, 2]
            z = t.zeros_like(exx)
            sig = t.stack([(lam + 2 * mu) * exx + lam * eyy, lam * exx + (lam + 2 * mu) * eyy,
                           lam * (exx + eyy), mu * gxy, z, z])                     # [6,2]
            hv = t.sqrt(0.5 * t.sum((P6 @ sig) ** 2, dim=0))
            obj = (hv * t.tensor(gh[i])).sum()
            g_ue, g_lam, g_mu = t.autograd.grad(obj, [ue, lam, mu], retain_graph=True, allow_unused=True)
            w = np.zeros(ndof)
            np.add.at(w, lmo, g_ue.numpy())
            w[2 * node_id - 2: 2 * node_id] += gy[i]
            psi = np.zeros(ndof)
            psi[free] = lu.solve(w[free])
            pe, uu = psi[self.lm], u[self.lm]
            gl = -np.einsum("ea,eab,eb->", pe, self._Kel, uu) + (float(g_lam) if g_lam is not None else 0.0)
            gm = -np.einsum("ea,eab,eb->", pe, self._Kem, uu) + (float(g_mu) if g_mu is not None else 0.0)
            (gxi,) = t.autograd.grad(gl * lam + gm * mu, xt)
            y[i] = u[2 * node_id - 2: 2 * node_id]
            h[i] = hv.detach().numpy()
            gx[i] = gxi.numpy()
        return y, h, gx


def read_mesh_text(text):
    """``read_mesh`` on an in-memory string."""
    import os
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
    try:
        return read_mesh(f.name)
    finally:
        os.unlink(f.name)
