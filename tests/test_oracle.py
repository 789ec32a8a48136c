"""CPU tests: the oracle against the reference's own outputs (golden vectors
produced by running the unmodified NumPy twin, tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest

import fem_oracle as fo
from conftest import relerr

TM, TS = (math.log(20.0), 0.0), (0.1, 0.015)


def test_mesh_generator_matches_reference_mesh(golden):
    m = fo.read_mesh_text(fo.cook_mesh_text(20, 10)) if hasattr(fo, "read_mesh_text") else None
    if m is None:
        import tempfile, os
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(fo.cook_mesh_text(20, 10))
        m = fo.read_mesh(f.name)
        os.unlink(f.name)
    assert np.array_equal(m["conn"], golden["IEN"])
    assert np.max(np.abs(m["coord"] - golden["coord"])) < 1e-12
    d = fo.assign_dof(m)
    assert np.array_equal(d["LM"], golden["LM"])
    assert np.array_equal(d["ID"], golden["ID"])
    assert np.array_equal(d["free_dof"], golden["free_dof"])
    assert np.array_equal(d["supp_dof"], golden["supp_dof"])
    assert np.max(np.abs(d["Pf"] - golden["Pf"])) < 1e-14  # reference file carries 1e-15 x-loads
    assert abs(d["Pf"].sum() - 50.0) < 1e-12


def test_shape_function_jacobian_pin(golden, oracle_mesh):
    mesh, dof = oracle_mesh
    xl = mesh["coord"][dof["IEN"][0] - 1, 1:3].T
    sg = fo.gauss_2x2()
    jac = [fo.shapef(sg[0:2, g], xl)[1] * sg[2, g] for g in range(4)]
    assert relerr(jac, golden["jac_ele1"]) < 1e-14


def test_loop_oracle_config1_fields(golden, oracle_mesh):
    """fem_test.py case (E=20, nu=0.3): every field of the reference run."""
    lo = fo.LoopOracle(*oracle_mesh)
    u, fint, strain, stress = lo.solve(20.0, 0.3)
    assert relerr(u, golden["c1_u"]) < 1e-11
    assert relerr(stress, golden["c1_stress"]) < 1e-11
    assert relerr(strain, golden["c1_strain"]) < 1e-11
    assert relerr(fint, golden["c1_Fint"]) < 1e-10
    assert relerr(fo.von_mises(stress[:, :, 11], (1, 3)), golden["c1_vm"]) < 1e-11
    assert abs(np.abs(u).sum() - 605.7948267813301) < 1e-8
    assert abs(np.abs(stress).sum() - 761.706958610109) < 1e-8


def test_loop_oracle_theta_samples(golden, oracle_mesh):
    lo = fo.LoopOracle(*oracle_mesh)
    y, h = fo.fem_fh_loop(lo, golden["x"][:3], TM, TS)
    assert relerr(y, golden["y"][:3]) < 1e-11
    assert relerr(h, golden["h"][:3]) < 1e-11


def test_torch_oracle_matches_reference(golden, torch_oracle):
    import torch
    x = torch.tensor(golden["x"])
    y, h = torch_oracle.fem_fh(x)
    u, _, stress = torch_oracle.fields(x)
    assert relerr(y.numpy(), golden["y"]) < 1e-11
    assert relerr(h.numpy(), golden["h"]) < 1e-11
    assert relerr(u.numpy(), golden["u"]) < 1e-11
    assert relerr(stress.numpy(), golden["stress"]) < 1e-11


def test_survey_known_answers(torch_oracle):
    """SURVEY.md 8(c) table (values printed by the reference NumPy twin)."""
    import torch
    y, h = torch_oracle.fem_fh(torch.tensor([[0.0, 0.0], [1.0, -1.0]], dtype=torch.float64))
    assert relerr(y[0].numpy(), [-4.218023950076504, 5.692883043125291]) < 1e-11
    assert relerr(h[1].numpy(), [0.2771999778329789, 0.2524508932200516]) < 1e-11


def test_torch_oracle_gradient_vs_finite_differences(torch_oracle):
    x = np.array([[1.0, -1.0], [-0.4, 2.0]])
    gy = np.array([[0.3, -0.7], [1.0, 0.2]])
    gh = np.array([[1.1, 0.4], [-0.5, 2.0]])
    _, _, gx = torch_oracle.vjp(x, gy, gh)
    import torch
    eps = 1e-5
    for i in range(2):
        for k in range(2):
            xp, xm = x.copy(), x.copy()
            xp[i, k] += eps
            xm[i, k] -= eps
            yp, hp = torch_oracle.fem_fh(torch.tensor(xp))
            ym, hm = torch_oracle.fem_fh(torch.tensor(xm))
            fd = ((yp - ym).numpy() * gy).sum() / (2 * eps) + ((hp - hm).numpy() * gh).sum() / (2 * eps)
            assert abs(fd - gx[i, k]) < 1e-7 * max(1.0, abs(gx[i, k]))
    # survey pin: at x=(1,-1), gy=(0.3,-0.7), gh=(1.1,0.4)
    assert abs(gx[0, 0] - 0.47551330398298) < 1e-10
    assert abs(gx[0, 1] - 0.00314673223922) < 1e-10


def test_elbo_broadcast_quirk(torch_oracle):
    """term2 averages over all B x (B*S) pairs (main_custom_training.py:205-214)."""
    import torch
    rng = np.random.default_rng(7)
    B, S = 3, 4
    mu = torch.tensor(rng.standard_normal((B, 2)) * 0.3)
    sig2 = torch.tensor(np.exp(rng.standard_normal((B, 2)) * 0.2))
    e = torch.tensor(rng.standard_normal((S, 2)))
    yb = torch.tensor(rng.standard_normal((B, 2)) + np.array([-4.2, 5.7]))
    loss, t1, t2, t3 = fo.elbo_step1_torch(torch_oracle, yb, mu, sig2, e, 0.1)
    theta = (e * sig2.sqrt().unsqueeze(1) + mu.unsqueeze(1)).reshape(-1, 2)
    f, _ = torch_oracle.fem_fh(theta)
    tot = sum(((yb[b] - f[j]) ** 2).sum() for b in range(B) for j in range(B * S))
    ref = -0.5 * 2 * math.log(2 * math.pi * 0.1) - 0.5 / 0.1 * tot / (B * B * S)
    assert abs(float(t2) - float(ref)) < 1e-12 * abs(float(ref))
    assert abs(float(loss) - float(t1 - t2 - t3)) < 1e-14


def test_sparse_oracle_vjp_equals_dense_autograd(oracle_mesh, torch_oracle):
    """SparseOracle.vjp (discrete adjoint on the SuperLU factor, used for the 80x40 parity tests) against
    TorchOracle.vjp (dense LU + full autograd) on the 20x10 mesh, incl. a supported observation node."""
    import fem_oracle as fo
    so = fo.SparseOracle(*oracle_mesh)
    rng = np.random.default_rng(21)
    x, gy, gh = rng.standard_normal((5, 2)), rng.standard_normal((5, 2)), rng.standard_normal((5, 2))
    y0, h0, g0 = torch_oracle.vjp(x, gy, gh)
    y1, h1, g1 = so.vjp(x, gy, gh, 231, 12)
    assert relerr(y1, y0) < 1e-11 and relerr(h1, h0) < 1e-11
    assert np.max(np.abs(g1 - g0) / np.maximum(np.abs(g0), 1e-300)) < 1e-8   # element-wise
    to2 = fo.TorchOracle(*oracle_mesh, node_id=23, ele_id=150, nipt_id=(1, 2))
    y0, h0, g0 = to2.vjp(x, gy, gh)
    y1, h1, g1 = so.vjp(x, gy, gh, 23, 150, (1, 2))
    assert relerr(g1, g0) < 1e-10


def test_oracle_jacobian_vs_reference_finite_differences(golden, torch_oracle):
    """Gradient pin on the reference's own code: the oracle's autograd Jacobian d(y, h)/dx against
    Richardson-extrapolated central differences of the UNMODIFIED reference NumPy twin
    (tests/golden/make_golden.py, arrays fd_x / fd_jac)."""
    for x, jref in zip(golden["fd_x"], golden["fd_jac"]):
        J = np.zeros((4, 2))
        for r in range(4):
            gy, gh = np.zeros((1, 2)), np.zeros((1, 2))
            (gy if r < 2 else gh)[0, r % 2] = 1.0
            J[r] = torch_oracle.vjp(x[None], gy, gh)[2][0]
        # dh/dx0 is exactly 0 (stresses do not depend on E under load control); FD noise there ~3e-12
        assert np.max(np.abs(J - jref)) < 2e-9 * np.max(np.abs(jref))
        big = np.abs(jref) > 1e-6
        assert np.max(np.abs(J - jref)[big] / np.abs(jref)[big]) < 2e-8


def test_loop_oracle_body_force_matches_reference(golden, oracle_mesh):
    """The part card's body force (src/mat_subroutine.py:113-116): the oracle's restatement against the
    unmodified reference twin run with body = (0.01, -0.02) (golden bf_*)."""
    lo = fo.LoopOracle(*oracle_mesh, body=(0.01, -0.02))
    u, fint, strain, stress = lo.solve(20.0, 0.3)
    assert relerr(u, golden["bf_u"]) < 1e-11 and relerr(stress, golden["bf_stress"]) < 1e-11
    assert relerr(fint, golden["bf_Fint"]) < 1e-10
    assert relerr(fo.von_mises(stress[:, :, 11], (1, 3)), golden["bf_vm"]) < 1e-11


def test_loop_oracle_plane_stress_branch(oracle_mesh):
    """Plane stress (src/mat_subroutine.py:283-290; not runnable on the reference itself under numpy >= 2):
    known properties of the restatement -- sigma_zz = 0, eps_33 = -v/(1-v)(eps_xx+eps_yy), softer than plane
    strain, equilibrium with the load."""
    ps = fo.LoopOracle(*oracle_mesh, stype=1)
    pe = fo.LoopOracle(*oracle_mesh, stype=2)
    u1, f1, e1, s1 = ps.solve(20.0, 0.3)
    u2, _, _, _ = pe.solve(20.0, 0.3)
    assert np.all(s1[2] == 0.0)
    assert relerr(e1[2], -0.3 / 0.7 * (e1[0] + e1[1])) < 1e-14
    assert np.abs(u1).sum() > 1.05 * np.abs(u2).sum()
    free = oracle_mesh[1]["free_dof"] - 1
    assert relerr(f1[free], oracle_mesh[1]["Pf"]) < 1e-9
