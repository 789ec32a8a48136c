"""CPU tests of the host-side logic and of the C-ABI library as an artefact
(loads, exports every declared symbol, refuses to run without a GPU)."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr


def test_feap_roundtrip_matches_reference_preprocessor(pkg, golden):
    """Product pre-processor on the generated 20x10 input vs the reference's
    pre-processor state (model_file.mat content captured in the golden file)."""
    P = pkg.PreProcessing
    md = P.modeldata_initialization_topopt(pkg.cook_membrane_feap(20, 10))
    assert md["mesh_info"]["nnodes"] == 231 and md["mesh_info"]["nele"] == 200
    assert np.max(np.abs(md["mesh_info"]["coord"] - golden["coord"])) < 1e-12
    di = md["dof_info"]
    for k in ("IEN", "LM", "ID", "free_dof", "supp_dof"):
        assert np.array_equal(di[k], golden[k]), k
    assert di["ndof"] == 462 and di["nfree"] == 440 and di["nsupp"] == 22
    assert np.max(np.abs(md["loading"]["Pf"].ravel() - golden["Pf"])) < 1e-14
    assert P.out_data["ele_stress"].shape == (6, 4, 200, 2)
    assert md["section"][0]["thk"] == 10 and md["material"][0]["E"] == 20.0


def test_feap_parser_handles_crlf_and_trailing_blocks(pkg):
    txt = pkg.cook_membrane_feap(4, 2).replace("\n", "\r\n") + "Parameters\r\nL = 48\r\nMATErial 1\r\nEND\r\n"
    m = pkg.fem_preprocess.parse_feap(txt)
    assert m["nnodes"] == 15 and m["nele"] == 8
    assert m["support"].shape == (3, 3) and m["nodal_load"].shape == (3, 3)
    assert abs(m["nodal_load"][:, 2].sum() - 50.0) < 1e-12


def test_feap_rejects_bad_input(pkg):
    with pytest.raises(ValueError):
        pkg.fem_preprocess.parse_feap("x\n 4 1 1 3 3 8\n\nCOORdinates ALL\n")
    with pytest.raises(ValueError):  # no loads
        txt = pkg.cook_membrane_feap(2, 2).split("FORCe conditions")[0] + "\nEND\n"
        pkg.PreProcessing.modeldata_initialization_topopt(txt)


def test_von_mises_host_matches_reference(pkg, golden):
    P = pkg.PreProcessing
    P.modeldata_initialization_topopt(pkg.cook_membrane_feap(20, 10))
    P.out_data["ele_stress"][:, :, :, 1] = golden["c1_stress"]
    vm = pkg.PostProcessing.von_mises_stress(2, 12, np.array([1, 3]))
    assert relerr(vm, golden["c1_vm"]) < 1e-14


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    hdr = open(os.path.join(ROOT, "include", "vbfem.h")).read()
    declared = set(re.findall(r"\b(vbfem_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg._lib.SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(rf"\bT {s}\b", out), s


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(pkg, golden_model):
    """Without a CUDA device the product path must fail loudly, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.VbfemError, match="no CUDA device|CPU fallback"):
        pkg.CookFemEngine(golden_model, device=0)


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "variational-bayesian-inference-for-computational-mechanics_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "fem_oracle" not in src and "/root/reference" not in src, f


def test_shard_range_partitions(pkg):
    for total in (0, 1, 7, 6400, 65536):
        for world in (1, 2, 3, 8):
            r = [pkg.elbo.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_term2_from_sums_equals_broadcast(pkg):
    import torch
    rng = np.random.default_rng(3)
    B, S = 5, 7
    f = torch.tensor(rng.standard_normal((B * S, 2)))
    yb = torch.tensor(rng.standard_normal((B, 2)))
    sums = torch.stack([f[:, 0].sum(), f[:, 1].sum(), (f ** 2).sum()])
    t2 = pkg.elbo.term2_from_sums(sums, yb, B * S, 0.1)
    l2 = -0.5 / 0.1 * ((yb.unsqueeze(1) - f) ** 2).sum(-1)  # [B, B*S], main_custom_training.py:210
    ref = -0.5 * 2 * math.log(2 * math.pi * 0.1) + l2.mean()
    assert abs(float(t2 - ref)) < 1e-13 * abs(float(ref))


def test_step1_model_shape_and_param_count(pkg):
    import torch
    m = pkg.elbo.make_step1_model()
    assert sum(p.numel() for p in m.parameters()) == 1884
    mu, sig, ls = m(torch.zeros(4, 2, dtype=torch.float64))
    assert mu.shape == (4, 2) and torch.allclose(sig, torch.exp(ls))


_WORKER = r'''
import os, sys, math
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import importlib
pkg = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")
import fem_oracle as fo
from fake_engine import OracleEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(11)
B, S = 3, 5
mu = torch.tensor(rng.standard_normal((B, 2)) * 0.3, requires_grad=True)
ls = torch.tensor(rng.standard_normal((B, 2)) * 0.2, requires_grad=True)
e = torch.tensor(rng.standard_normal((S, 2)))
yb = torch.tensor(rng.standard_normal((B, 2)) * 0.5 + np.array([-4.2, 5.7]))
eng = OracleEngine()
loss_fn = pkg.elbo.Step1Loss(eng, e, 0.1, group=None, rank=rank, world=world)
loss = loss_fn(yb, mu, torch.exp(ls), ls)
loss.backward()
# single-process oracle of the same thing
mu2 = mu.detach().clone().requires_grad_(True); ls2 = ls.detach().clone().requires_grad_(True)
ref, *_ = fo.elbo_step1_torch(eng.oracle, yb, mu2, torch.exp(ls2), e, 0.1)
ref.backward()
err = max(abs(float(loss - ref)) / abs(float(ref)),
          float((mu.grad - mu2.grad).abs().max() / mu2.grad.abs().max()),
          float((ls.grad - ls2.grad).abs().max() / ls2.grad.abs().max()))
print("RANK", rank, "ERR", err, flush=True)
assert err < 1e-10, err
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2])
def test_step1_loss_sharded_over_gloo_ranks(tmp_path, world):
    """World-size-2 gloo run of the sharded ELBO step: every rank owns a slice
    of the Monte-Carlo samples, one all-reduce, result equals the unsharded
    oracle (value and gradients)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", OMP_NUM_THREADS="2")
    procs = []
    for r in range(world):
        procs.append(subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), WORLD_SIZE=str(world)),
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ERR" in o


def test_step1_loss_library_side_form_equals_framework_form(pkg):
    """Step1Loss with the whole loss from the engine (elbo_step1_loss: KL terms and loss value next to the data term)
    against the form that keeps term1 / term3 in the framework: same value, same gradients through exp(log_sig2) --
    the autograd glue of elbo.Step1Fused on the CPU, with the oracle as the engine."""
    import torch
    from fake_engine import OracleEngine, OracleEngineFused
    rng = np.random.default_rng(17)
    B, S = 3, 4
    e = torch.tensor(rng.standard_normal((S, 2)))
    yb = torch.tensor(rng.standard_normal((B, 2)) * 0.5 + np.array([-4.2, 5.7]))
    res = []
    for eng, fused in ((OracleEngineFused(), True), (OracleEngine(), True), (OracleEngineFused(), False)):
        mu = torch.tensor(rng.standard_normal((B, 2)) * 0.0 + np.linspace(-0.3, 0.3, 2 * B).reshape(B, 2), requires_grad=True)
        ls = torch.tensor(np.linspace(-0.4, 0.2, 2 * B).reshape(B, 2), requires_grad=True)
        loss_fn = pkg.elbo.Step1Loss(eng, e, 0.1, fused=fused)
        loss = 2.5 * loss_fn(yb, mu, torch.exp(ls), ls)      # a non-unit cotangent reaches the saved gradients
        loss.backward()
        res.append((float(loss), mu.grad.clone(), ls.grad.clone()))
    for other in res[1:]:
        assert abs(res[0][0] - other[0]) < 1e-12 * abs(other[0])
        assert float((res[0][1] - other[1]).abs().max()) < 1e-12 * float(other[1].abs().max())
        assert float((res[0][2] - other[2]).abs().max()) < 1e-12 * float(other[2].abs().max())


def test_step2_loss_equals_oracle_broadcast(pkg):
    """Step2Loss (sufficient statistics of h) against the statement-by-statement restatement of
    main_custom_training.py:338-384 with its [B, B*S] broadcast: value and gradients w.r.t. the z nets."""
    import torch
    import fem_oracle as fo
    from fake_engine import OracleEngine
    rng = np.random.default_rng(21)
    B, S = 3, 4
    eng = OracleEngine()
    mu = torch.tensor(rng.standard_normal((B, 2)) * 0.3)
    sig2 = torch.tensor(np.exp(rng.standard_normal((B, 2)) * 0.2))
    e = torch.tensor(rng.standard_normal((S, 2)))
    post_m = torch.tensor(rng.standard_normal((B, 2)) * 0.1 - 1.3)
    post_s = torch.tensor(np.abs(rng.standard_normal((B, 2))) * 0.05)
    vals = []
    for impl in range(2):
        zm = torch.tensor(np.full((B, 2), -1.3) + 0.05 * np.arange(B)[:, None], requires_grad=True)
        lzs = torch.tensor(np.full((B, 2), -3.0) + 0.1 * np.arange(2)[None, :], requires_grad=True)
        zs = torch.exp(lzs)
        if impl == 0:
            loss = pkg.elbo.Step2Loss(eng, e, 3e-3, alpha=0.7)(mu, sig2, zm, zs, lzs, post_m, post_s)
        else:
            loss, *_ = fo.elbo_step2_torch(eng.oracle, mu, sig2, zm, zs, lzs, post_m, post_s, e, 3e-3, alpha=0.7)
        loss.backward()
        vals.append((float(loss), zm.grad.clone(), lzs.grad.clone()))
    assert abs(vals[0][0] - vals[1][0]) < 1e-12 * abs(vals[1][0])
    assert float((vals[0][1] - vals[1][1]).abs().max()) < 1e-11 * float(vals[1][1].abs().max())
    assert float((vals[0][2] - vals[1][2]).abs().max()) < 1e-11 * float(vals[1][2].abs().max())


_WORKER2 = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import importlib
pkg = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")
import fem_oracle as fo
from fake_engine import OracleEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(31)
B, S = 3, 5
mu = torch.tensor(rng.standard_normal((B, 2)) * 0.3); sig2 = torch.tensor(np.exp(rng.standard_normal((B, 2)) * 0.2))
e = torch.tensor(rng.standard_normal((S, 2)))
zm = torch.tensor(rng.standard_normal((B, 2)) * 0.1 - 1.3, requires_grad=True)
lzs = torch.tensor(rng.standard_normal((B, 2)) * 0.1 - 3.0, requires_grad=True)
pm = torch.tensor(rng.standard_normal((B, 2)) * 0.1 - 1.3); ps = torch.tensor(np.abs(rng.standard_normal((B, 2))) * 0.05)
eng = OracleEngine()
loss = pkg.elbo.Step2Loss(eng, e, 3e-3, alpha=1.0, rank=rank, world=world)(mu, sig2, zm, torch.exp(lzs), lzs, pm, ps)
loss.backward()
zm2 = zm.detach().clone().requires_grad_(True); lzs2 = lzs.detach().clone().requires_grad_(True)
ref, *_ = fo.elbo_step2_torch(eng.oracle, mu, sig2, zm2, torch.exp(lzs2), lzs2, pm, ps, e, 3e-3)
ref.backward()
err = max(abs(float(loss - ref)) / abs(float(ref)),
          float((zm.grad - zm2.grad).abs().max() / zm2.grad.abs().max()),
          float((lzs.grad - lzs2.grad).abs().max() / lzs2.grad.abs().max()))
print("RANK", rank, "ERR", err, flush=True)
assert err < 1e-10, err
dist.destroy_process_group()
'''


def test_step2_loss_sharded_over_gloo_ranks(tmp_path):
    script = tmp_path / "worker2.py"
    script.write_text(_WORKER2.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29612", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), WORLD_SIZE="2"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ERR" in o


def test_plan_layout_on_host(pkg, golden_model, monkeypatch):
    """vbfem_plan: the kernel choice / numbering / orientation / front split vbfem_create would make, computed
    without a GPU.  Cook 20x10 with the reference's observation set-up: short-side numbering (b = 25 instead of
    43); the warp-per-sample kernel by default, the on-chip two-front kernel when it is disabled or cannot take
    the observation set-up (middle block on the observed element, observed node in the bottom front)."""
    plan = pkg.fem_solver.plan_layout(golden_model)
    assert plan["kernel_variant"] == 4 and plan["nfree"] == 440 and plan["half_bw"] == 25
    # second generation of the warp kernel: the (K_lam, K_mu) band table [440][26] x 16 B in shared memory + per warp
    # 640 B (last panel's Minv^T, 1/d, flag) + 1088 B (small vectors) + the 40-row window of u and four adjoints
    # (forward / fused-adjoint launches: sixteen warps, one adjoint vector; Jacobian mode: twelve warps, four)
    assert plan["smem_bytes"] == 440 * 26 * 16 + 16 * (640 + 1088 + 40 * 2 * 8)
    monkeypatch.setenv("VBFEM_WARP2_NW16", "0")
    assert pkg.fem_solver.plan_layout(golden_model)["smem_bytes"] == 440 * 26 * 16 + 12 * (640 + 1088 + 40 * 5 * 8)
    monkeypatch.delenv("VBFEM_WARP2_NW16")
    # first generation (element matrices per sample, gather table): what a smaller shared memory falls back to
    per_warp = 640 + 4 * 512 + (44 * 36 + 2) * 8     # last panel's Minv^T / 1/d / flag, staging area, ring of 44 element matrices
    monkeypatch.setenv("VBFEM_WARP_V1", "1")
    assert pkg.fem_solver.plan_layout(golden_model)["smem_bytes"] == 12 * per_warp + 3816 * 8 + 56 * 8   # twelve warps + the packed gather table + row table
    monkeypatch.delenv("VBFEM_WARP_V1")
    # a supported observed node has no unit vectors: warp kernel; a node in the middle of the band order: front kernel
    assert pkg.fem_solver.plan_layout(golden_model, node_id=1, ele_id=50)["kernel_variant"] == 4
    assert pkg.fem_solver.plan_layout(golden_model, node_id=1, ele_id=60)["kernel_variant"] == 4
    assert pkg.fem_solver.plan_layout(golden_model, node_id=21, ele_id=12)["kernel_variant"] == 2
    # less shared memory: eight warps instead of twelve
    assert pkg.fem_solver.plan_layout(golden_model, smem_per_sm=190000)["smem_bytes"] == 8 * per_warp + 3816 * 8 + 56 * 8
    monkeypatch.setenv("VBFEM_WARP", "0")
    plan = pkg.fem_solver.plan_layout(golden_model)
    assert plan == {"kernel_variant": 2, "nfree": 440, "half_bw": 25, "twist_row": 220, "bottom_cols": 194,
                    "flipped": 0, "smem_bytes": 109888}
    # observed node ahead of the observed element: the band order is reversed
    p2 = pkg.fem_solver.plan_layout(golden_model, node_id=23, ele_id=150)
    assert p2["kernel_variant"] == 2 and p2["flipped"] == 1
    assert p2["twist_row"] + 26 + p2["bottom_cols"] == 440 and p2["twist_row"] >= 32 and p2["bottom_cols"] >= 32
    # observed node inside the observed element's rows: generic kernel; element too close to the end of the band
    # for two fronts: blocked panel kernel
    assert pkg.fem_solver.plan_layout(golden_model, node_id=116, ele_id=110)["kernel_variant"] == 0
    assert pkg.fem_solver.plan_layout(golden_model, node_id=1, ele_id=60)["kernel_variant"] == 3
    # supported observed node: no unit vectors, still the front kernel
    assert pkg.fem_solver.plan_layout(golden_model, node_id=1, ele_id=50)["kernel_variant"] == 2
    # a smaller shared memory does not fit two samples per SM
    assert pkg.fem_solver.plan_layout(golden_model, smem_per_sm=200000)["kernel_variant"] == 3


@pytest.mark.parametrize("nx,ny,variant,b,variant_no_warp", [(24, 8, 4, 21, 2), (20, 9, 4, 23, 2), (16, 8, 4, 21, 3),
                                                            (30, 10, 4, 25, 3), (80, 40, 3, 85, 3), (40, 20, 3, 45, 3)])
def test_plan_layout_other_meshes(pkg, nx, ny, variant, b, variant_no_warp, monkeypatch):
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(nx, ny))
    plan = pkg.fem_solver.plan_layout(md, node_id=(nx + 1) * (ny + 1), ele_id=nx // 2 + 2)
    assert plan["kernel_variant"] == variant and plan["half_bw"] == b
    assert plan["nfree"] == 2 * nx * (ny + 1)
    monkeypatch.setenv("VBFEM_WARP", "0")
    assert pkg.fem_solver.plan_layout(md, node_id=(nx + 1) * (ny + 1), ele_id=nx // 2 + 2)["kernel_variant"] == variant_no_warp


def test_h5io_roundtrip_and_layout(pkg, tmp_path):
    """h5io: write -> read round trip of a save_data-shaped dictionary (float64, stored transposed like
    hdf5storage's MATLAB-compatible mode), and the structural fields a libhdf5 reader checks first."""
    import struct
    rng = np.random.default_rng(0)
    d = {"y_data": rng.standard_normal((50, 2)), "y_scaled_data": rng.standard_normal((50, 2)),
         "z_data": rng.standard_normal((50, 2)), "log_z_data": rng.standard_normal((50, 2)),
         "z_scaled_data": rng.standard_normal((50, 2)), "y_mean": rng.standard_normal((1, 2)),
         "y_std": rng.standard_normal((1, 2)), "z_mean": rng.standard_normal((1, 2)),
         "z_std": rng.standard_normal((1, 2)), "e_data": rng.standard_normal((7, 2))}
    f = str(tmp_path / "data.h5")
    pkg.h5io.write(data=d, filename=f)
    back = pkg.h5io.read(filename=f)
    assert sorted(back) == sorted(d)
    for k in d:
        assert back[k].dtype == np.float64 and np.array_equal(back[k], d[k])
    raw = open(f, "rb").read()
    assert raw.startswith(b"MATLAB 7.3 MAT-file") and raw[512:520] == b"\x89HDF\r\n\x1a\n" and raw[520] == 0
    base, _, eof, _ = struct.unpack_from("<QQQQ", raw, 512 + 24)
    assert base == 512 and base + eof == len(raw)
    untransposed = pkg.h5io.read(filename=f, matlab_compatible=False)
    assert untransposed["y_data"].shape == (2, 50)      # MATLAB order on disk
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not an hdf5 file" * 100)
        pkg.h5io.read(filename=str(bad))


def test_h5io_reads_a_deflate_shuffle_fletcher32_file(pkg):
    """The committed fixture tests/golden/chunked_filters.h5 was cut from the reference's shipped
    data_fem_test_big_noise.h5 layout by tests/golden/make_h5_fixture.py: chunked datasets with
    shuffle + deflate + fletcher32, version-1 B-trees.  Values are the recipe's seeded arrays."""
    f = os.path.join(ROOT, "tests", "golden", "chunked_filters.h5")
    d = pkg.h5io.read(filename=f)
    rng = np.random.default_rng(7)
    a = rng.standard_normal((300, 2))
    b = rng.standard_normal((1, 2))
    assert np.array_equal(d["a_data"], a) and np.array_equal(d["b_mean"], b)


def test_generate_data_fem_is_seed_compatible(pkg, monkeypatch):
    """generate_data_fem consumes NumPy's global generator in upstream's order (theta, err, eta, e_data;
    src/data_generation_2sam_more_loss.py:64-73), the constructor draws nothing, save_data writes the ten
    datasets of save_data upstream (src/data_generation_2sam_more_loss.py:256-268)."""
    M = pkg.MeasurementData
    calls = []

    def fake_fem(x):
        calls.append(np.array(x))
        return [np.stack([x[:, 0], 2 * x[:, 1]], 1), np.exp(0.1 * x) + 5.0]
    monkeypatch.setattr(M, "fem_fh_fun_loop_rev", staticmethod(fake_fem))
    np.random.seed(123)
    md = M(n_sam=6, ne_sam=3, d_y=2, d_z=2, d_theta=2, sig_e=0.1, sig_eta=3e-3)
    assert md.e_data.shape == (6, 2) and not md.e_data.any()     # upstream's initial shapes, no draws
    md.generate_data_fem()
    np.random.seed(123)
    theta = np.random.randn(6, 2)
    err = np.sqrt(0.1) * np.random.randn(6, 2)
    eta = np.sqrt(3e-3) * np.random.randn(6, 2)
    e_data = np.random.randn(3, 2)
    assert np.array_equal(calls[0], theta) and np.array_equal(md.e_data, e_data)
    assert np.array_equal(md.y_data, np.stack([theta[:, 0], 2 * theta[:, 1]], 1) + err)
    assert np.array_equal(md.z_data, np.exp(0.1 * theta) + 5.0 + eta)
    assert np.array_equal(md.log_z_data, np.log(md.z_data)) and md.y_mean.shape == (1, 2)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "d.h5")
        md.save_data("/", f)
        back = M.load_data("/", f)
    assert sorted(back) == sorted(["y_data", "y_scaled_data", "z_data", "log_z_data", "z_scaled_data", "y_mean",
                                   "y_std", "z_mean", "z_std", "e_data"])
    assert np.array_equal(back["y_data"], md.y_data) and np.array_equal(back["e_data"], md.e_data)


def test_oversized_mesh_fails_cleanly(pkg):
    """A mesh whose band rows do not fit 16 bits (n > 32767 free dofs) is refused with an error code and a message
    (it used to double-free the half-built handle); garbage connectivity is refused too."""
    nx, ny = 130, 130                                   # 2 * 130 * 131 = 34060 free dofs
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(nx, ny))
    with pytest.raises(pkg.VbfemError, match="too large"):
        pkg.fem_solver.plan_layout(md, node_id=(nx + 1) * (ny + 1), ele_id=12)
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(4, 4))
    md["dof_info"]["IEN"] = md["dof_info"]["IEN"].copy()
    md["dof_info"]["IEN"][0, 0] = 999
    with pytest.raises(pkg.VbfemError, match="out of range"):
        pkg.fem_solver.plan_layout(md, node_id=25, ele_id=2)
