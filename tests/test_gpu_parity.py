"""GPU parity tests: the CUDA path, called through the C ABI, against (i) the
reference's own outputs (golden vectors of the unmodified NumPy twin), (ii) the
CPU oracle on the same seeded inputs, (iii) size-independent properties at the
benchmark's full size.  Tolerance: 1e-9 relative, float64 (BASELINE.json
north_star); most checks hold to ~1e-12."""
import importlib
import math

import numpy as np
import pytest

from conftest import model_from_golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _t(a, eng):
    import torch
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=eng.device)


def conftest_pkg_name():
    import conftest
    return conftest.PKG


def golden_model_of(engine):
    import conftest
    return conftest.model_from_golden(np.load(conftest.os.path.join(conftest.ROOT, "tests", "golden", "ref_numpy_twin.npz")))


def test_native_library_is_loaded(engine, pkg):
    maps = open("/proc/self/maps").read()
    assert "libvbfem.so" in maps
    assert engine.info["nfree"] == 440 and engine.info["half_bw"] == 25  # short-side numbering
    assert engine.info["kernel_variant"] == 4 and engine.info["block_threads"] == 512   # warp-per-sample kernel, 16 warps
    # the host-only plan (vbfem_plan, unit-tested without a GPU) is what vbfem_create built
    plan = pkg.fem_solver.plan_layout(golden_model_of(engine))
    for k in ("kernel_variant", "nfree", "half_bw", "twist_row", "smem_bytes"):
        assert plan[k] == engine.info[k], (k, plan, engine.info)


def test_forward_matches_reference_golden(engine, golden):
    y, h = engine.forward(_t(golden["x"], engine))
    assert relerr(y.cpu().numpy(), golden["y"]) < TOL
    assert relerr(h.cpu().numpy(), golden["h"]) < TOL
    assert engine.status(len(golden["x"]))[0] == 0


def test_fields_config1_matches_fem_test(engine, golden):
    """fem_test.py case (cards E=20, nu=0.3): u, stress, strain, F_int."""
    out = engine.fields(emat=_t([[20.0, 0.3]], engine))
    assert relerr(out["u"][0].cpu().numpy(), golden["c1_u"]) < TOL
    assert relerr(out["stress"][0].cpu().numpy(), golden["c1_stress"]) < TOL
    assert relerr(out["strain"][0].cpu().numpy(), golden["c1_strain"]) < TOL
    assert relerr(out["fint"][0].cpu().numpy(), golden["c1_Fint"]) < 1e-8  # F_int_f ~ Pf, reactions
    u = out["u"][0].cpu().numpy()
    assert abs(np.abs(u).sum() - 605.7948267813301) < 1e-7


def test_fields_theta_samples(engine, golden):
    out = engine.fields(x=_t(golden["x"], engine), want=("u", "stress"))
    assert relerr(out["u"].cpu().numpy(), golden["u"]) < TOL
    assert relerr(out["stress"].cpu().numpy(), golden["stress"]) < TOL


def test_fea_solution_drop_in(pkg, golden):
    """fem_test.py flow on the new backend: initialise from the mesh text, call
    FemSolver.fea_solution, read results where upstream's scripts read them."""
    P = pkg.PreProcessing
    P.modeldata_initialization_topopt(pkg.cook_membrane_feap(20, 10))
    pkg.FemSolver.fea_solution(input_data=None)
    step_id = len(P.out_data["step"])
    assert step_id == 2
    assert relerr(P.sol_data["u_n1"].ravel(), golden["c1_u"]) < TOL
    assert relerr(P.out_data["step"][1]["nodal_disp"], golden["c1_nodal_disp"]) < TOL
    vm = pkg.PostProcessing.von_mises_stress(step_id, 12, np.array([1, 3]))
    assert relerr(vm, golden["c1_vm"]) < TOL
    assert P.out_data["step"][1]["tol_vec"][0] < 1e-9  # energy-norm check of src/fem_solver.py:106-124


def _errs(a, b):
    """(norm-wise, element-wise) relative errors; element-wise with a floor of 1e-3 of the largest entry."""
    a, b = np.asarray(a), np.asarray(b)
    return relerr(a, b), float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * np.abs(b).max())))


def test_forward_adjoint_vs_oracle_full_batch(engine, torch_oracle):
    """Config 2 at its full size: ALL 4096 seeded samples (the bench's own inputs: x seed 0, cotangents
    seed 1 split into distinct gy / gh) against the oracle, norm-wise and element-wise."""
    import bench
    x, gy, gh = bench.inputs(0)
    assert not np.array_equal(gy, gh)
    yo, ho, gxo = [np.concatenate(p) for p in zip(*[torch_oracle.vjp(x[i:i + 256], gy[i:i + 256], gh[i:i + 256])
                                                    for i in range(0, len(x), 256)])]
    y, h, gx = engine.forward_backward(_t(x, engine), _t(gy, engine), _t(gh, engine))
    assert engine.status(len(x))[0] == 0
    y, h, g = y.cpu().numpy(), h.cpu().numpy(), gx.cpu().numpy()
    for name, a, b in (("y", y, yo), ("h", h, ho), ("gx", g, gxo)):
        nw, ew = _errs(a, b)
        print(f"config 2, 4096 samples, {name}: norm-wise {nw:.2e}, element-wise {ew:.2e}")
        assert nw < TOL and ew < TOL
    # each gradient component against its own magnitude per sample (gx1 ~ 3e-3 is 150x smaller than gx0)
    for k in range(2):
        ck = np.max(np.abs(g[:, k] - gxo[:, k])) / np.max(np.abs(gxo[:, k]))
        print(f"config 2, gx[:, {k}]: component-wise {ck:.2e}")
        assert ck < TOL
    # split forward(keep) + backward gives the same numbers as the fused launch
    y2, h2 = engine.forward(_t(x, engine), keep_factor=True)
    gx2 = engine.backward(_t(gy, engine), _t(gh, engine))
    assert relerr(y2.cpu().numpy(), y) < 1e-13
    assert relerr(gx2.cpu().numpy(), g) < 1e-11


def test_front_kernel_full_batch_vs_warp_kernel(pkg, engine, golden_model, monkeypatch):
    """The on-chip two-front kernel (VBFEM_WARP=0) and the warp-per-sample kernel are independent
    implementations of the same solve: all 4096 benchmark samples agree to 1e-11."""
    import bench
    monkeypatch.setenv("VBFEM_WARP", "0")
    front = pkg.CookFemEngine(golden_model, device=0)
    assert front.info["kernel_variant"] == 2 and engine.info["kernel_variant"] == 4
    x, gy, gh = (_t(a, engine) for a in bench.inputs(0))
    a = engine.forward_backward(x, gy, gh)
    b = front.forward_backward(x, gy, gh)
    for u, v in zip(a, b):
        assert relerr(u.cpu().numpy(), v.cpu().numpy()) < 1e-11
    ja, jb = engine.forward_jac(x)[2], front.forward_jac(x)[2]
    assert relerr(ja.cpu().numpy(), jb.cpu().numpy()) < 1e-11
    front.close()


def test_warp_kernel_generations_agree(pkg, engine, golden_model, torch_oracle, monkeypatch):
    """Second generation of the warp kernel (K = lambda K_lam + mu K_mu from the band table, contraction on the same
    table) against the first (VBFEM_WARP_V1=1: element matrices per sample, element-wise contraction) on all 4096
    benchmark samples, in all three modes, with sixteen and with twelve warps per SM; the oracle on a slice."""
    import bench
    assert engine.info["kernel_variant"] == 4 and engine.info["panel_ring"] == 0      # the default is generation 2
    monkeypatch.setenv("VBFEM_WARP_V1", "1")
    v1 = pkg.CookFemEngine(golden_model, device=0)
    monkeypatch.delenv("VBFEM_WARP_V1")
    monkeypatch.setenv("VBFEM_WARP2_NW16", "0")
    v2_12 = pkg.CookFemEngine(golden_model, device=0)
    assert v1.info["panel_ring"] == 44 and v2_12.info["panel_ring"] == 0
    x, gy, gh = (_t(a, engine) for a in bench.inputs(0))
    ref = v1.forward_backward(x, gy, gh)
    jref = v1.forward_jac(x)[2]
    for eng in (engine, v2_12):
        for u, v in zip(eng.forward_backward(x, gy, gh), ref):
            assert relerr(u.cpu().numpy(), v.cpu().numpy()) < 1e-11
        for u, v in zip(eng.forward(x), ref[:2]):
            assert relerr(u.cpu().numpy(), v.cpu().numpy()) < 1e-11
        assert relerr(eng.forward_jac(x)[2].cpu().numpy(), jref.cpu().numpy()) < 1e-11
        assert eng.status(x.shape[0])[0] == 0
    xs, gys, ghs = (a[:96] for a in bench.inputs(0))
    yo, ho, gxo = torch_oracle.vjp(xs, gys, ghs)
    y, h, gx = v2_12.forward_backward(_t(xs, engine), _t(gys, engine), _t(ghs, engine))
    assert max(relerr(y.cpu().numpy(), yo), relerr(h.cpu().numpy(), ho), relerr(gx.cpu().numpy(), gxo)) < TOL
    v1.close()
    v2_12.close()


def test_jacobian_vs_reference_finite_differences(engine, golden):
    """Gradient pinned to the reference's OWN code: d(y, h)/dx from the CUDA Jacobian mode against
    central differences of the unmodified reference NumPy twin (golden fd_x / fd_jac)."""
    _, _, jac = engine.forward_jac(_t(golden["fd_x"], engine))
    J, jref = jac.cpu().numpy(), golden["fd_jac"]
    assert np.max(np.abs(J - jref)) < 2e-9 * np.max(np.abs(jref))
    big = np.abs(jref) > 1e-6
    assert np.max(np.abs(J - jref)[big] / np.abs(jref)[big]) < 2e-8


def test_kept_jacobians_are_ticketed(pkg, engine):
    """Stale-state hazards of the differentiable drop-in: backward never applies another call's
    Jacobians (ticket), a plain launch in between does not disturb the kept ones, and autograd nodes
    own their state (two FEM calls in one graph)."""
    import torch
    rng = np.random.default_rng(12)
    xa, xb = _t(rng.standard_normal((40, 2)), engine), _t(rng.standard_normal((40, 2)), engine)
    gy, gh = _t(rng.standard_normal((40, 2)), engine), _t(rng.standard_normal((40, 2)), engine)
    _, _, ga = engine.forward_backward(xa, gy, gh)
    _, _, gb = engine.forward_backward(xb, gy, gh)
    engine.forward(xa, keep_factor=True)
    ta = engine.keep_ticket()
    engine.forward(xb)                                  # plain launch in between: kept Jacobians untouched
    engine.forward_backward(xb, gy, gh)
    assert torch.equal(engine.backward(gy, gh, ticket=ta), engine.backward(gy, gh))
    assert relerr(engine.backward(gy, gh, ticket=ta).cpu().numpy(), ga.cpu().numpy()) < 1e-11
    engine.forward(xb, keep_factor=True)                # a second keep: the first ticket is stale now
    tb = engine.keep_ticket()
    assert tb != ta
    with pytest.raises(pkg.VbfemError):
        engine.backward(gy, gh, ticket=ta)
    assert relerr(engine.backward(gy, gh, ticket=tb).cpu().numpy(), gb.cpu().numpy()) < 1e-11
    with pytest.raises(pkg.VbfemError):                  # wrong batch size
        engine.backward(gy[:7].contiguous(), gh[:7].contiguous())
    # two differentiable calls in one autograd graph, backward in one sweep
    M = pkg.MeasurementData
    pkg.PreProcessing.reset()
    pkg.PreProcessing.model_data = golden_model_of(engine)
    M.theta_mean, M.theta_std = np.array([math.log(20.0), 0.0]), np.array([0.1, 0.015])
    M.node_id, M.ele_id, M.nipt_id = 231, 12, np.array([1, 3], dtype=int)
    x1 = xa.clone().requires_grad_(True)
    x2 = xb.clone().requires_grad_(True)
    y1, h1 = M.fem_fh_fun_loop_rev(x1)
    y2, h2 = M.fem_fh_fun_loop_rev(x2)
    ((y1 * gy).sum() + (h1 * gh).sum() + (y2 * gy).sum() + (h2 * gh).sum()).backward()
    assert relerr(x1.grad.cpu().numpy(), ga.cpu().numpy()) < 1e-11
    assert relerr(x2.grad.cpu().numpy(), gb.cpu().numpy()) < 1e-11


def test_survey_gradient_pin(engine):
    y, h, gx = engine.forward_backward(_t([[1.0, -1.0]], engine), _t([[0.3, -0.7]], engine),
                                       _t([[1.1, 0.4]], engine))
    g = gx.cpu().numpy()[0]
    assert abs(g[0] - 0.47551330398298) < 1e-10 and abs(g[1] - 0.00314673223922) < 1e-10


def test_full_batch_properties(engine):
    """N=4096 (benchmark size): u ~ 1/E, stresses independent of E under load
    control, status clean, gradient consistent with finite differences of the
    CUDA forward itself."""
    n = 4096
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 2))
    y, h = engine.forward(_t(x, engine))
    assert engine.status(n)[0] == 0
    xs = x.copy()
    d = rng.standard_normal(n)
    xs[:, 0] += d
    y2, h2 = engine.forward(_t(xs, engine))
    y, h, y2, h2 = (a.cpu().numpy() for a in (y, h, y2, h2))
    assert relerr(y2, y * np.exp(-0.1 * d)[:, None]) < 1e-10
    assert relerr(h2, h) < 1e-10
    gy, gh = rng.standard_normal((n, 2)), rng.standard_normal((n, 2))
    _, _, gx = engine.forward_backward(_t(x, engine), _t(gy, engine), _t(gh, engine))
    eps = 1e-5
    for k in range(2):
        xp, xm = x.copy(), x.copy()
        xp[:, k] += eps
        xm[:, k] -= eps
        yp, hp = engine.forward(_t(xp, engine))
        ym, hm = engine.forward(_t(xm, engine))
        fd = (((yp - ym).cpu().numpy() * gy).sum(1) + ((hp - hm).cpu().numpy() * gh).sum(1)) / (2 * eps)
        g = gx.cpu().numpy()[:, k]
        assert np.max(np.abs(fd - g)) < 2e-6 * max(1.0, np.abs(g).max())


def test_edge_cases(engine, torch_oracle):
    import torch
    # empty batch
    y, h = engine.forward(torch.empty(0, 2, dtype=torch.float64, device=engine.device))
    assert y.shape == (0, 2) and h.shape == (0, 2)
    # single sample and ragged sizes around the resident-CTA count
    resident = engine.info["num_sms"] * engine.info["ctas_per_sm"]
    for n in (1, 3, resident - 1, resident + 1):
        x = np.random.default_rng(n).standard_normal((n, 2))
        y, h = engine.forward(_t(x, engine))
        idx = np.unique(np.array([0, n // 2, n - 1]))
        yo, ho = torch_oracle.fem_fh(torch.tensor(x[idx]))
        assert relerr(y.cpu().numpy()[idx], yo.numpy()) < TOL
        assert relerr(h.cpu().numpy()[idx], ho.numpy()) < TOL
    # nearly incompressible (nu -> 0.4999): still positive definite, still accurate to 1e-7
    x = np.array([[0.0, 500.0], [2.0, -500.0]])
    y, h = engine.forward(_t(x, engine))
    yo, ho = torch_oracle.fem_fh(torch.tensor(x))
    assert relerr(y.cpu().numpy(), yo.numpy()) < 1e-7
    assert engine.status(2)[0] == 0
    # nu == 0.5 exactly (lambda = inf): flagged, not silently wrong
    y, h = engine.forward(_t([[0.0, 1e6]], engine))
    assert engine.status(1)[0] == 1


def test_host_entry_points_equal_device_entry_points(engine):
    n = 300
    rng = np.random.default_rng(5)
    x, gy, gh = rng.standard_normal((n, 2)), rng.standard_normal((n, 2)), rng.standard_normal((n, 2))
    y, h, gx = engine.forward_backward(_t(x, engine), _t(gy, engine), _t(gh, engine))
    yh, hh, gxh = engine.forward_backward_host(x, gy, gh)
    assert np.array_equal(yh, y.cpu().numpy()) and np.array_equal(hh, h.cpu().numpy())
    assert np.array_equal(gxh, gx.cpu().numpy())
    yf, hf = engine.forward_host(x)
    assert np.array_equal(yf, yh) and np.array_equal(hf, hh)


def test_measurement_data_drop_in_autograd(pkg, golden, golden_model, torch_oracle):
    """MeasurementData.fem_fh_fun_loop_rev as the differentiable operator
    (the role it plays at main_custom_training.py:191-196)."""
    import torch
    P, M = pkg.PreProcessing, pkg.MeasurementData
    P.reset()
    P.model_data = golden_model
    M.theta_mean, M.theta_std = np.array([math.log(20.0), 0.0]), np.array([0.1, 0.015])
    M.node_id, M.ele_id, M.nipt_id = 231, 12, np.array([1, 3], dtype=int)
    dev = torch.device("cuda", 0)
    x = torch.tensor(golden["x"], device=dev, requires_grad=True)
    y, h = M.fem_fh_fun_loop_rev(x)
    w = torch.linspace(0.5, 1.5, 32, dtype=torch.float64, device=dev).reshape(16, 2)
    ((y * w).sum() + (h * w.flip(0)).sum()).backward()
    _, _, gxo = torch_oracle.vjp(golden["x"], w.cpu().numpy(), w.flip(0).cpu().numpy())
    assert relerr(y.detach().cpu().numpy(), golden["y"]) < TOL
    assert relerr(x.grad.cpu().numpy(), gxo) < TOL
    # NumPy in -> NumPy out (eager callers: generate_data_fem, postprocess_lib)
    yn, hn = M.fem_fh_fun_loop_rev(golden["x"])
    assert isinstance(yn, np.ndarray) and relerr(hn, golden["h"]) < TOL
    assert relerr(M.fem_f_fun(golden["x"][1]), golden["y"][1]) < TOL
    assert relerr(M.fem_h_fun(golden["x"][1]), golden["h"][1]) < TOL


def test_elbo_step1_fused_vs_oracle(pkg, engine, torch_oracle):
    import torch
    import fem_oracle as fo
    rng = np.random.default_rng(9)
    B, S = 4, 6
    mu = rng.standard_normal((B, 2)) * 0.3
    ls = rng.standard_normal((B, 2)) * 0.2
    e = rng.standard_normal((S, 2))
    yb = rng.standard_normal((B, 2)) * 0.5 + np.array([-4.2, 5.7])
    # oracle
    mu_o = torch.tensor(mu, requires_grad=True)
    ls_o = torch.tensor(ls, requires_grad=True)
    ref, *_ = fo.elbo_step1_torch(torch_oracle, torch.tensor(yb), mu_o, torch.exp(ls_o), torch.tensor(e), 0.1)
    ref.backward()
    # CUDA, unsharded
    mu_c = _t(mu, engine).requires_grad_(True)
    ls_c = _t(ls, engine).requires_grad_(True)
    loss_fn = pkg.elbo.Step1Loss(engine, _t(e, engine), 0.1)
    loss = loss_fn(_t(yb, engine), mu_c, torch.exp(ls_c), ls_c)
    loss.backward()
    assert abs(float(loss) - float(ref)) < TOL * abs(float(ref))
    assert relerr(mu_c.grad.cpu().numpy(), mu_o.grad.numpy()) < TOL
    assert relerr(ls_c.grad.cpu().numpy(), ls_o.grad.numpy()) < TOL
    # the same with the KL terms and the loss value in the framework (fused=False): both forms agree to rounding
    mu_f = _t(mu, engine).requires_grad_(True)
    ls_f = _t(ls, engine).requires_grad_(True)
    loss_f = pkg.elbo.Step1Loss(engine, _t(e, engine), 0.1, fused=False)(_t(yb, engine), mu_f, torch.exp(ls_f), ls_f)
    loss_f.backward()
    assert abs(float(loss) - float(loss_f)) < 1e-13 * abs(float(loss_f))
    assert relerr(mu_c.grad.cpu().numpy(), mu_f.grad.cpu().numpy()) < 1e-13
    assert relerr(ls_c.grad.cpu().numpy(), ls_f.grad.cpu().numpy()) < 1e-13
    # shard-count invariance: partials of 1, 2, 3, 8 shards add up to the same numbers
    full = engine.elbo_step1_partials(mu_c.detach(), torch.exp(ls_c.detach()), _t(e, engine), _t(yb, engine), 0.1)
    for world in (2, 3, 8):
        acc = [torch.zeros_like(t) for t in full[:3]]
        for r in range(world):
            lo, hi = pkg.elbo.shard_range(B * S, r, world)
            part = engine.elbo_step1_partials(mu_c.detach(), torch.exp(ls_c.detach()), _t(e, engine),
                                              _t(yb, engine), 0.1, lo, hi)
            for a, p in zip(acc, part[:3]):
                a += p
        for a, f in zip(acc, full[:3]):
            assert relerr(a.cpu().numpy(), f.cpu().numpy()) < 1e-12


def test_elbo_step1_config3_size_vs_oracle(pkg, engine, torch_oracle):
    """Config 3 at its real size (B = 64, S = 100: 6400 solves with adjoint per step) against the oracle's
    statement-by-statement ELBO (incl. the [B, B*S] broadcast) differentiated by torch autograd."""
    import torch
    import fem_oracle as fo
    rng = np.random.default_rng(3)
    B, S = 64, 100
    e = rng.standard_normal((S, 2))
    yb = np.random.default_rng(2).standard_normal((B, 2)) * np.array([0.53, 0.65]) + np.array([-4.24, 5.71])
    mu = rng.standard_normal((B, 2)) * 0.5
    ls = rng.standard_normal((B, 2)) * 0.3 - 0.5
    mu_o = torch.tensor(mu, requires_grad=True)
    ls_o = torch.tensor(ls, requires_grad=True)
    ref, *_ = fo.elbo_step1_torch(torch_oracle, torch.tensor(yb), mu_o, torch.exp(ls_o), torch.tensor(e), 0.1)
    ref.backward()
    mu_c = _t(mu, engine).requires_grad_(True)
    ls_c = _t(ls, engine).requires_grad_(True)
    loss = pkg.elbo.Step1Loss(engine, _t(e, engine), 0.1)(_t(yb, engine), mu_c, torch.exp(ls_c), ls_c)
    loss.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < TOL * abs(float(ref.detach()))
    for name, a, b in (("d/dmu", mu_c.grad.cpu().numpy(), mu_o.grad.numpy()),
                       ("d/dlog_sig", ls_c.grad.cpu().numpy(), ls_o.grad.numpy())):
        nw, ew = _errs(a, b)
        print(f"config 3 (B=64, S=100) ELBO gradient {name}: norm-wise {nw:.2e}, element-wise {ew:.2e}")
        assert nw < TOL and ew < 1e-8


def test_refined_mesh_80x40_batch_vs_oracle(pkg):
    """Config 4: 64 samples of the 80x40 mesh, forward AND adjoint against the sparse CPU oracle
    (its own discrete adjoint on the SuperLU factor), fused and Jacobian mode; panel-kernel variant."""
    import fem_oracle as fo
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
    eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
    assert eng.info["kernel_variant"] == 3 and eng.info["panel_blocks"] == 11
    m = fo.read_mesh_text(fo.cook_mesh_text(80, 40))
    so = fo.SparseOracle(m, fo.assign_dof(m))
    n = 64
    x = np.random.default_rng(4).standard_normal((n, 2))
    gy = np.random.default_rng(5).standard_normal((n, 2))
    gh = np.random.default_rng(6).standard_normal((n, 2))
    yo, ho, gxo = so.vjp(x, gy, gh, 3321, 12)
    y, h, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    assert eng.status(n)[0] == 0
    for name, a, b in (("y", y, yo), ("h", h, ho), ("gx", gx, gxo)):
        nw, ew = _errs(a.cpu().numpy(), b)
        print(f"config 4 (80x40), 64 samples, {name}: norm-wise {nw:.2e}, element-wise {ew:.2e}")
        assert nw < TOL and ew < 1e-8
    y2, h2, jac = eng.forward_jac(_t(x, eng))
    gx2 = eng.jac_vjp(jac, _t(gy, eng), _t(gh, eng))
    assert relerr(y2.cpu().numpy(), yo) < TOL and relerr(gx2.cpu().numpy(), gxo) < TOL
    y3, h3 = eng.forward(_t(x, eng))
    assert relerr(y3.cpu().numpy(), yo) < TOL and relerr(h3.cpu().numpy(), ho) < TOL
    # batch sizes around the resident-CTA count (persistent CTAs take a second sample)
    resident = eng.info["num_sms"] * eng.info["ctas_per_sm"]
    xr = np.random.default_rng(40).standard_normal((resident + 3, 2))
    yr, hr = eng.forward(_t(xr, eng))
    idx = [0, resident - 1, resident, resident + 2]
    yq, hq = so.fem_fh(xr[idx], 3321, 12)
    assert relerr(yr.cpu().numpy()[idx], yq) < TOL and relerr(hr.cpu().numpy()[idx], hq) < TOL
    eng.close()


def test_generic_kernel_forced(pkg, golden_model, torch_oracle, monkeypatch):
    """The generic per-column kernel (the fallback every mesh can take, and the fields path) still serves
    forward / fused adjoint / Jacobian mode when neither fast kernel is allowed."""
    monkeypatch.setenv("VBFEM_FORCE_GENERIC", "1")
    eng = pkg.CookFemEngine(golden_model, device=0)
    assert eng.info["kernel_variant"] == 0
    n = 48
    x = np.random.default_rng(31).standard_normal((n, 2))
    gy = np.random.default_rng(32).standard_normal((n, 2))
    gh = np.random.default_rng(33).standard_normal((n, 2))
    yo, ho, gxo = torch_oracle.vjp(x, gy, gh)
    y, h, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    assert relerr(y.cpu().numpy(), yo) < TOL and relerr(h.cpu().numpy(), ho) < TOL
    assert relerr(gx.cpu().numpy(), gxo) < TOL
    eng.forward(_t(x, eng), keep_factor=True)
    assert relerr(eng.backward(_t(gy, eng), _t(gh, eng)).cpu().numpy(), gxo) < TOL
    eng.close()


def test_refined_mesh_80x40(pkg):
    """Config 4: Cook 80x40 through the same entry points (band spills to HBM)
    against the sparse CPU oracle; gradient against central differences."""
    import fem_oracle as fo
    txt = pkg.cook_membrane_feap(80, 40)
    P = pkg.PreProcessing
    md = P.modeldata_initialization_topopt(txt)
    eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
    assert eng.info["nfree"] == 6560 and eng.info["half_bw"] == 85
    m = fo.read_mesh_text(fo.cook_mesh_text(80, 40))
    so = fo.SparseOracle(m, fo.assign_dof(m))
    x = np.random.default_rng(4).standard_normal((6, 2))
    yo, ho = so.fem_fh(x, 3321, 12)
    y, h = eng.forward(_t(x, eng))
    assert relerr(y.cpu().numpy(), yo) < TOL and relerr(h.cpu().numpy(), ho) < TOL
    gy, gh = np.ones((6, 2)), np.full((6, 2), 0.5)
    _, _, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    eps = 1e-5
    for k in range(2):
        xp, xm = x.copy(), x.copy()
        xp[:, k] += eps
        xm[:, k] -= eps
        yp, hp = so.fem_fh(xp, 3321, 12)
        ym, hm = so.fem_fh(xm, 3321, 12)
        fd = (((yp - ym) * gy).sum(1) + ((hp - hm) * gh).sum(1)) / (2 * eps)
        assert np.max(np.abs(fd - gx.cpu().numpy()[:, k])) < 1e-6 * max(1.0, np.abs(fd).max())
    # forward(keep) + backward (the autograd / tf.custom_gradient route: factor kept in HBM) == fused
    y2, h2 = eng.forward(_t(x, eng), keep_factor=True)
    gx2 = eng.backward(_t(gy, eng), _t(gh, eng))
    assert relerr(y2.cpu().numpy(), yo) < TOL and relerr(h2.cpu().numpy(), ho) < TOL
    assert relerr(gx2.cpu().numpy(), gx.cpu().numpy()) < TOL
    eng.close()


@pytest.mark.parametrize("node_id,ele_id,nipt_id,variant,warp", [
    (231, 12, (1, 3), 4, "1"),    # the reference's set-up: observed node at the end of the band order -> warp kernel
    (231, 12, (1, 3), 2, "0"),    # the same on the on-chip two-front kernel: node behind the element in band order
    (21, 12, (2, 4), 2, "1"),     # tip of the bottom edge (middle of the band order), other Gauss points: front kernel
    (23, 150, (1, 2), 2, "1"),    # node AHEAD of the element: the band order is flipped internally
    (1, 50, (3, 4), 4, "1"),      # supported node: y == 0, gradient through h only
    (1, 50, (3, 4), 2, "0"),
    (116, 110, (1, 3), 0, "1"),   # node of the observed element itself: generic kernel
    (1, 60, (3, 4), 4, "1"),
    (1, 60, (3, 4), 3, "0"),      # element too close to the end of the band for two fronts: blocked panel kernel
])
def test_other_observation_setups(pkg, golden_model, oracle_mesh, node_id, ele_id, nipt_id, variant, warp, monkeypatch):
    """Every kernel's layout (orientation, middle block, unit vectors, right-hand-side rows) is derived from the
    observation set-up; every choice must agree with the oracle."""
    import fem_oracle as fo
    monkeypatch.setenv("VBFEM_WARP", warp)
    eng = pkg.CookFemEngine(golden_model, device=0, node_id=node_id, ele_id=ele_id, nipt_id=nipt_id)
    assert eng.info["kernel_variant"] == variant
    to = fo.TorchOracle(*oracle_mesh, node_id=node_id, ele_id=ele_id, nipt_id=nipt_id)
    n = 64
    x = np.random.default_rng(7).standard_normal((n, 2))
    gy = np.random.default_rng(8).standard_normal((n, 2))
    gh = np.random.default_rng(9).standard_normal((n, 2))
    yo, ho, gxo = to.vjp(x, gy, gh)
    y, h, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    scale = max(float(np.abs(yo).max()), 1e-30)
    assert float(np.abs(y.cpu().numpy() - yo).max()) < TOL * max(scale, 1.0)
    assert relerr(h.cpu().numpy(), ho) < TOL
    assert relerr(gx.cpu().numpy(), gxo) < TOL
    y2, h2 = eng.forward(_t(x, eng), keep_factor=True)       # Jacobian mode + J^T g
    gx2 = eng.backward(_t(gy, eng), _t(gh, eng))
    assert float(np.abs(y2.cpu().numpy() - yo).max()) < TOL * max(scale, 1.0)
    assert relerr(h2.cpu().numpy(), ho) < TOL
    assert relerr(gx2.cpu().numpy(), gxo) < TOL
    y3, h3 = eng.forward(_t(x, eng))                          # forward only
    assert float(np.abs(y3.cpu().numpy() - yo).max()) < TOL * max(scale, 1.0)
    assert relerr(h3.cpu().numpy(), ho) < TOL
    assert eng.status(n)[0] == 0
    eng.close()


def test_graphed_elbo_step_equals_eager(pkg, engine):
    """The CUDA-graph replay of a training step (nets + fused FEM op + Adam) follows the eager
    step exactly: same losses for the same batches and initial weights."""
    import torch
    dev = engine.device
    rng = np.random.default_rng(11)
    yd = rng.standard_normal((6, 16, 2)) * np.array([0.53, 0.65]) + np.array([-4.24, 5.71])
    e_data = _t(np.random.default_rng(3).standard_normal((20, 2)), engine)
    losses = {}
    for graph in (False, True):
        model = pkg.elbo.make_step1_model(device=dev, seed=5)
        opt = pkg.elbo.make_step1_optimizer_capturable(model)
        step = pkg.elbo.GraphedStep1(model, opt, pkg.elbo.Step1Loss(engine, e_data, 0.1), 16, dev, use_graph=graph)
        seq = []
        if not graph:                      # the graphed trainer warms up with 3 real steps on its first batch
            for _ in range(3):
                step.step(yd[0])
        for b in yd:
            seq.append(float(step.step(b)))
        assert step.graphed == graph
        losses[graph] = np.array(seq)
    assert np.max(np.abs(losses[True] - losses[False])) < 1e-10 * np.max(np.abs(losses[False]))
    # pipelined stepping (the host runs ahead of the GPU, losses are read afterwards): the same sequence
    model = pkg.elbo.make_step1_model(device=dev, seed=5)
    step = pkg.elbo.GraphedStep1(model, pkg.elbo.make_step1_optimizer_capturable(model),
                                 pkg.elbo.Step1Loss(engine, e_data, 0.1), 16, dev)
    tickets = [step.step_async(b, depth=3) for b in yd[:3]]
    first = [step.loss_of(t) for t in tickets]
    tickets = [step.step_async(b, depth=3) for b in yd[3:]]
    seq = np.array(first + [step.loss_of(t) for t in tickets])
    assert np.max(np.abs(seq - losses[True])) < 1e-10 * np.max(np.abs(losses[True]))
    with pytest.raises(ValueError):
        step.loss_of(0)        # its staging slot has been reused


def test_elbo_step2_fused_vs_oracle(pkg, engine, torch_oracle):
    """Step 2 (main_custom_training.py:304-384) on the fused forward-only op: the logz pre-pass and
    the loss (term4 - term5) * alpha + add_loss with gradients w.r.t. the z nets."""
    import torch
    import fem_oracle as fo
    rng = np.random.default_rng(41)
    B, S = 8, 25
    mu = rng.standard_normal((B, 2)) * 0.3
    sig2 = np.exp(rng.standard_normal((B, 2)) * 0.2)
    e = rng.standard_normal((S, 2))
    eta = math.sqrt(3e-3) * rng.standard_normal((S, 2))
    # pre-pass: moments of log z over the reparameterised samples (main_custom_training.py:311-328)
    pm, ps = pkg.elbo.logz_posterior_moments(engine, _t(mu, engine), _t(sig2, engine), _t(e, engine), _t(eta, engine))
    theta = (e[None] * np.sqrt(sig2)[:, None] + mu[:, None]).reshape(-1, 2)
    _, ho = torch_oracle.fem_fh(torch.tensor(theta))
    logz = np.log(ho.numpy().reshape(B, S, 2) + eta[None])
    assert relerr(pm.cpu().numpy(), logz.mean(1)) < TOL and relerr(ps.cpu().numpy(), logz.var(1)) < 1e-8
    res = []
    for impl in range(2):
        dev = engine.device if impl == 0 else torch.device("cpu")
        zm = torch.tensor(logz.mean(1) + 0.02, device=dev, requires_grad=True)
        lzs = torch.tensor(np.log(logz.var(1)) + 0.1, device=dev, requires_grad=True)
        T = lambda a: torch.tensor(a, device=dev)
        if impl == 0:
            loss = pkg.elbo.Step2Loss(engine, T(e), 3e-3, alpha=0.5)(T(mu), T(sig2), zm, torch.exp(lzs), lzs,
                                                                       T(logz.mean(1)), T(logz.var(1)))
        else:
            loss, *_ = fo.elbo_step2_torch(torch_oracle, T(mu), T(sig2), zm, torch.exp(lzs), lzs, T(logz.mean(1)),
                                           T(logz.var(1)), T(e), 3e-3, alpha=0.5)
        loss.backward()
        res.append((float(loss), zm.grad.cpu().numpy(), lzs.grad.cpu().numpy()))
    assert abs(res[0][0] - res[1][0]) < TOL * abs(res[1][0])
    assert relerr(res[0][1], res[1][1]) < TOL and relerr(res[0][2], res[1][2]) < TOL


@pytest.mark.parametrize("nx,ny,variant,warp", [
    (24, 8, 4, "1"),    # n = 432, narrower band (b = 21 < 25): warp kernel, window padded to three blocks
    (24, 8, 2, "0"),    #          front kernel with a zero-padded band
    (20, 9, 4, "1"),    # n = 400, b = 23
    (20, 9, 2, "0"),    #          other front lengths
    (16, 8, 4, "1"),    # n = 288
    (14, 8, 4, "1"),    # n = 252 = 4 mod 8: four leading pad rows (pivot mu in the band table of the warp kernel)
    (10, 8, 4, "1"),    # n = 180 = 4 mod 8, 23 panels
    (16, 8, 3, "0"),    #          too small for the front kernel's shared-memory layout -> blocked panel kernel
    (30, 10, 4, "1"),   # n = 660: the larger gather table leaves room for eight warps
    (30, 10, 3, "0"),   #          the band no longer fits twice per SM -> blocked panel kernel
    (40, 20, 3, "1"),   # n = 1680, b = 45: panel kernel with a 6-block window
])
def test_other_mesh_sizes(pkg, nx, ny, variant, warp, monkeypatch):
    """Cook membranes of other sizes through the same entry points (mesh text -> preprocessor ->
    engine), against the oracle built from the same text: forward, fused adjoint, Jacobian mode."""
    import torch
    import fem_oracle as fo
    monkeypatch.setenv("VBFEM_WARP", warp)
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(nx, ny))
    node_id, ele_id = (nx + 1) * (ny + 1), nx // 2 + 2
    eng = pkg.CookFemEngine(md, device=0, node_id=node_id, ele_id=ele_id)
    assert eng.info["kernel_variant"] == variant, eng.info
    m = fo.read_mesh_text(fo.cook_mesh_text(nx, ny))
    dof = fo.assign_dof(m)
    to = fo.TorchOracle(m, dof, node_id=node_id, ele_id=ele_id)
    n = 40
    x = np.random.default_rng(nx).standard_normal((n, 2))
    gy = np.random.default_rng(ny).standard_normal((n, 2))
    gh = np.random.default_rng(nx + ny).standard_normal((n, 2))
    yo, ho, gxo = to.vjp(x, gy, gh)
    y, h, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    assert relerr(y.cpu().numpy(), yo) < TOL and relerr(h.cpu().numpy(), ho) < TOL
    assert relerr(gx.cpu().numpy(), gxo) < TOL
    y2, h2 = eng.forward(_t(x, eng), keep_factor=True)
    gx2 = eng.backward(_t(gy, eng), _t(gh, eng))
    assert relerr(y2.cpu().numpy(), yo) < TOL and relerr(h2.cpu().numpy(), ho) < TOL
    assert relerr(gx2.cpu().numpy(), gxo) < TOL
    assert eng.status(n)[0] == 0
    eng.close()


def test_tf_bridge_contract_with_stand_in_tf(pkg, engine, golden, torch_oracle, monkeypatch):
    """tf_bridge.make_fem_fh_op / install driven end to end against tests/fake_tf.py (TensorFlow is not in
    this image): forward through the py_function island and the DLPack round trip, the custom_gradient
    contract (forward -> grad fn -> backward), two outstanding calls whose gradients are taken out of order."""
    import sys
    import types
    import fake_tf
    tf = fake_tf.make_module()
    monkeypatch.setitem(sys.modules, "tensorflow", tf)
    bridge = importlib.import_module(conftest_pkg_name() + ".tf_bridge")
    M = pkg.MeasurementData
    pkg.PreProcessing.reset()
    pkg.PreProcessing.model_data = golden_model_of(engine)
    ref_dg = types.SimpleNamespace(MeasurementData=types.SimpleNamespace(
        theta_mean=np.array([math.log(20.0), 0.0]), theta_std=np.array([0.1, 0.015]), node_id=231, ele_id=12,
        nipt_id=np.array([1, 3], dtype=int), fem_fh_fun_loop_rev=None))
    op = bridge.install(ref_dg)
    fn = ref_dg.MeasurementData.fem_fh_fun_loop_rev
    xa, xb = tf.constant(golden["x"][:8]), tf.constant(golden["x"][8:])
    ya, ha = fn(xa)
    grad_a = op.last_grad
    yb, hb = fn(xb)
    grad_b = op.last_grad
    assert ya.shape == (8, 2) and relerr(ya.numpy(), golden["y"][:8]) < TOL and relerr(hb.numpy(), golden["h"][8:]) < TOL
    rng = np.random.default_rng(17)
    gy, gh = rng.standard_normal((8, 2)), rng.standard_normal((8, 2))
    gxb = grad_b(tf.constant(gy), tf.constant(gh))       # out of order: b first, then a
    gxa = grad_a(tf.constant(gy), tf.constant(gh))
    assert relerr(gxa.numpy(), torch_oracle.vjp(golden["x"][:8], gy, gh)[2]) < TOL
    assert relerr(gxb.numpy(), torch_oracle.vjp(golden["x"][8:], gy, gh)[2]) < TOL
    assert tuple(M.nipt_id) == (1, 3) and M.node_id == 231   # install keeps the class attributes in sync


def test_small_batch_host_path_and_mcmc_log_posterior(pkg, engine, golden, torch_oracle):
    """The one-sample-at-a-time callers (src/postprocess_lib.py:78-103): batches of 1, 8 and 64 on mapped
    pinned memory give the device path's numbers bit for bit; logp_y_2d equals upstream's expression on the
    oracle's f; a short Metropolis chain runs through it; the large-batch KDE sampler agrees with the oracle."""
    import torch
    for n in (1, 8, 64, 65):
        x = np.random.default_rng(n).standard_normal((n, 2))
        gy, gh = np.random.default_rng(n + 1).standard_normal((n, 2)), np.random.default_rng(n + 2).standard_normal((n, 2))
        y, h, gx = engine.forward_backward(_t(x, engine), _t(gy, engine), _t(gh, engine))
        yh, hh, gxh = engine.forward_backward_host(x, gy, gh)
        yf, hf = engine.forward_host(x)
        assert np.array_equal(yh, y.cpu().numpy()) and np.array_equal(gxh, gx.cpu().numpy())
        assert np.array_equal(yf, yh) and np.array_equal(hf, hh) and np.array_equal(hh, h.cpu().numpy())
    M, PP = pkg.MeasurementData, pkg.postprocess_lib.PostProcess
    pkg.PreProcessing.reset()
    pkg.PreProcessing.model_data = golden_model_of(engine)
    M.theta_mean, M.theta_std = np.array([math.log(20.0), 0.0]), np.array([0.1, 0.015])
    M.node_id, M.ele_id, M.nipt_id = 231, 12, np.array([1, 3], dtype=int)
    y_obs, sig_e = np.array([-4.1, 5.6]), 0.1
    logp = PP.logp_y_2d(y_obs, sig_e)
    for th in golden["x"][:4]:
        f = torch_oracle.fem_fh(torch.tensor(th[None]))[0].numpy()
        ref = -0.5 / sig_e * np.sum((y_obs - f) ** 2) - np.log(2 * np.pi * sig_e) - 0.5 * np.sum(th ** 2) - np.log(2 * np.pi)
        assert abs(logp(th) - ref) < 1e-9 * abs(ref)
    chain = pkg.postprocess_lib.metropolis_chain(logp, np.zeros(2), 60, burn=20, thin=2, scale=0.5,
                                                 rng=np.random.default_rng(0))
    assert chain.shape == (20, 2) and np.isfinite(chain).all() and len(np.unique(chain[:, 0])) > 3
    zm, zv = PP.moments_2d_case4_method1(np.zeros((3, 2)), np.ones((3, 2)), 3e-3, 50, rng=np.random.default_rng(1))
    rng = np.random.default_rng(1)
    theta = rng.standard_normal((50, 2))            # std = 1, mean = 0: the same draw for every observation
    eta = np.sqrt(3e-3) * rng.standard_normal((50, 2))
    z = torch_oracle.fem_fh(torch.tensor(theta))[1].numpy() + eta
    assert relerr(zm, np.tile(z.mean(0), (3, 1))) < TOL and relerr(zv, np.tile(z.var(0), (3, 1))) < 1e-8


def test_generality_plane_stress_body_force_elementwise(pkg, golden, golden_model, oracle_mesh):
    """SURVEY 8(f) row 4 -- the card options the reference stubs out, on the generic kernel: plane stress
    (section stype 1), a body force (part body; pinned to the reference twin's own run), one (E, nu) per
    element, and the Newton-Raphson solver option."""
    import copy
    import torch
    import fem_oracle as fo
    # ---- body force through FemSolver.fea_solution with the part card set, against the reference golden
    P = pkg.PreProcessing
    P.modeldata_initialization_topopt(pkg.cook_membrane_feap(20, 10))
    P.model_data["part"][0]["body"] = np.array([[0.01], [-0.02], [0.0]])
    P.model_data["solution_control"]["solver"] = 2      # Newton-Raphson option: converges in the first iteration
    pkg.FemSolver.fea_solution(input_data=None)
    assert relerr(P.sol_data["u_n1"].ravel(), golden["bf_u"]) < TOL
    assert relerr(P.out_data["ele_stress"][:, :, :, 1], golden["bf_stress"]) < TOL
    assert relerr(P.sol_data["F_int"].ravel(), golden["bf_Fint"]) < 1e-8
    assert relerr(pkg.PostProcessing.von_mises_stress(2, 12, np.array([1, 3])), golden["bf_vm"]) < TOL
    assert int(P.out_data["step"][1]["iter_vec"][0, 0]) == 1
    # the batched theta path with the same body force: fast kernel (the load vector is all that changes)
    eng = pkg.CookFemEngine(P.model_data, device=0)
    assert eng.info["kernel_variant"] == 4
    lo = fo.LoopOracle(*oracle_mesh, body=(0.01, -0.02))
    x = golden["x"][:3]
    yo, ho = fo.fem_fh_loop(lo, x, (math.log(20.0), 0.0), (0.1, 0.015))
    y, h = eng.forward(_t(x, eng))
    assert relerr(y.cpu().numpy(), yo) < TOL and relerr(h.cpu().numpy(), ho) < TOL
    eng.close()
    # ---- plane stress: forward, fields and the adjoint (finite differences of the oracle), generic kernel
    eng = pkg.CookFemEngine(golden_model, device=0, stype=1)
    assert eng.info["kernel_variant"] == 0
    lo = fo.LoopOracle(*oracle_mesh, stype=1)
    u, fint, strain, stress = lo.solve(20.0, 0.3)
    out = eng.fields(emat=_t([[20.0, 0.3]], eng))
    assert relerr(out["u"][0].cpu().numpy(), u) < TOL and relerr(out["stress"][0].cpu().numpy(), stress) < TOL
    assert relerr(out["strain"][0].cpu().numpy(), strain) < TOL
    x = golden["x"][:4]
    yo, ho = fo.fem_fh_loop(lo, x, (math.log(20.0), 0.0), (0.1, 0.015))
    gy, gh = np.random.default_rng(51).standard_normal((4, 2)), np.random.default_rng(52).standard_normal((4, 2))
    y, h, gx = eng.forward_backward(_t(x, eng), _t(gy, eng), _t(gh, eng))
    assert relerr(y.cpu().numpy(), yo) < TOL and relerr(h.cpu().numpy(), ho) < TOL
    eps = 1e-5
    for k in range(2):
        xp, xm = x.copy(), x.copy()
        xp[:, k] += eps
        xm[:, k] -= eps
        yp, hp = eng.forward(_t(xp, eng))
        ym, hm = eng.forward(_t(xm, eng))
        fd = (((yp - ym).cpu().numpy() * gy).sum(1) + ((hp - hm).cpu().numpy() * gh).sum(1)) / (2 * eps)
        assert np.max(np.abs(fd - gx.cpu().numpy()[:, k])) < 2e-6 * max(1.0, np.abs(fd).max())
    eng.close()
    # ---- one (E, nu) per element
    eng = pkg.CookFemEngine(golden_model, device=0)
    rng = np.random.default_rng(53)
    Ee, ve = 20.0 * np.exp(0.2 * rng.standard_normal(200)), 0.2 + 0.2 * rng.random(200)
    lo = fo.LoopOracle(*oracle_mesh)
    u, fint, strain, stress = lo.solve(Ee, ve)
    emat = _t(np.stack([Ee, ve], 1)[None], eng)
    out = eng.fields_elementwise(emat)
    assert relerr(out["u"][0].cpu().numpy(), u) < TOL and relerr(out["stress"][0].cpu().numpy(), stress) < TOL
    assert relerr(out["fint"][0].cpu().numpy(), fint) < 1e-8
    assert relerr(out["y"][0].cpu().numpy(), u[460:462]) < TOL
    assert relerr(out["h"][0].cpu().numpy(), fo.von_mises(stress[:, :, 11], (1, 3))) < TOL
    eng.close()
