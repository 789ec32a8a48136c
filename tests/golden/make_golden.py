"""Generate golden vectors by running the UNMODIFIED reference NumPy twin.

Runs only in the build container (needs /root/reference).  It imports the
reference's own modules (src/fem_preprocess.py, src/fem_solver.py,
src/mat_subroutine.py, src/fem_postprocess.py,
src/data_generation_2sam_more_loss.py) behind import shims for the packages
that are absent here (tensorflow, hdf5storage, h5py, matplotlib) and records

  * the pre-processor state (coord, IEN, LM, ID, free/supp dof, Pf) that pins
    the mesh/DOF interface (fem_preprocess.py:114-443),
  * config 1 (fem_test.py: E=20, nu=0.3): full u, sigma, eps, von Mises; the same with a body force in the part card (bf_*),
  * theta-parameterised solves through MeasurementData.fem_f_fun / fem_h_fun
    (data_generation_2sam_more_loss.py:98-125) for a list of seeded x,
  * finite-difference Jacobians d(y, h)/dx of those same reference functions (gradient pins).

Output: tests/golden/ref_numpy_twin.npz  (committed; this script is the recipe).
Usage:  python tests/golden/make_golden.py
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _shim():
    for name in ("hdf5storage", "h5py", "matplotlib", "matplotlib.pyplot", "tensorflow"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    tf = sys.modules["tensorflow"]
    tf.function = lambda f=None, **kw: (f if f is not None else (lambda g: g))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))


def main():
    _shim()
    work = tempfile.mkdtemp(prefix="vbfem_golden_")
    shutil.copy(os.path.join(REF, "Armero_cooksm_20x10.txt"), work)
    os.chdir(work)
    import fem_preprocess as fp
    import fem_solver as fs
    import fem_postprocess as fpp
    from src import data_generation_2sam_more_loss as dg

    fp.PreProcessing.modeldata_initialization_topopt("Armero_cooksm_20x10.txt", "model_file.mat")
    md = fp.PreProcessing.model_data
    out = {}
    out["coord"] = np.asarray(md["mesh_info"]["coord"], dtype=np.float64)
    out["IEN"] = np.asarray(md["dof_info"]["IEN"], dtype=np.int64)
    out["LM"] = np.asarray(md["dof_info"]["LM"], dtype=np.int64)
    out["ID"] = np.asarray(md["dof_info"]["ID"], dtype=np.int64)
    out["free_dof"] = np.asarray(md["dof_info"]["free_dof"], dtype=np.int64)
    out["supp_dof"] = np.asarray(md["dof_info"]["supp_dof"], dtype=np.int64)
    out["Pf"] = np.asarray(md["loading"]["Pf"].toarray(), dtype=np.float64).ravel()
    out["loc_i_head"] = np.asarray(md["dof_info"]["loc_i_array"][:128], dtype=np.int64)
    out["loc_j_head"] = np.asarray(md["dof_info"]["loc_j_array"][:128], dtype=np.int64)
    out["jac_ele1"] = np.asarray(fp.PreProcessing.topo_data["element_kdata"]["jac"], dtype=np.float64)

    # ---- config 1: fem_test.py (cards default E=20, nu=0.3) -----------------
    fs.FemSolver.fea_solution(input_data=None)
    od, sd = fp.PreProcessing.out_data, fp.PreProcessing.sol_data
    out["c1_u"] = np.asarray(sd["u_n1"].toarray(), dtype=np.float64).ravel()
    out["c1_Fint"] = np.asarray(sd["F_int"].toarray(), dtype=np.float64).ravel()
    out["c1_stress"] = np.asarray(od["ele_stress"][:, :, :, 1], dtype=np.float64).copy()
    out["c1_strain"] = np.asarray(od["ele_strain"][:, :, :, 1], dtype=np.float64).copy()
    out["c1_nodal_disp"] = np.asarray(od["step"][1]["nodal_disp"], dtype=np.float64)
    out["c1_vm"] = np.asarray(fpp.PostProcessing.von_mises_stress(2, 12, np.array([1, 3])), dtype=np.float64)
    out["c1_tol"] = np.asarray(od["step"][1]["tol_vec"], dtype=np.float64)

    # ---- the cards' other branches, still on the unmodified reference: plane stress (section stype = 1,
    #      src/mat_subroutine.py:283-290) and a constant body force (part body, src/mat_subroutine.py:113-116)
    def solve_variant(tag):
        fp.PreProcessing.out_data["step"] = fp.PreProcessing.out_data["step"][:1]
        fs.FemSolver.fea_solution(input_data=None)
        out[tag + "_u"] = np.asarray(sd["u_n1"].toarray(), dtype=np.float64).ravel()
        out[tag + "_Fint"] = np.asarray(sd["F_int"].toarray(), dtype=np.float64).ravel()
        out[tag + "_stress"] = np.asarray(od["ele_stress"][:, :, :, 1], dtype=np.float64).copy()
        out[tag + "_strain"] = np.asarray(od["ele_strain"][:, :, :, 1], dtype=np.float64).copy()
        out[tag + "_vm"] = np.asarray(fpp.PostProcessing.von_mises_stress(2, 12, np.array([1, 3])), dtype=np.float64)

    # (the plane-stress branch itself cannot be run unmodified under numpy >= 2: src/mat_subroutine.py:52 assigns the
    #  size-1 array eps33 to a scalar slot and raises; it is pinned through the oracle's restatement only)
    md["part"][0]["body"] = np.array([[0.01], [-0.02], [0.0]])
    solve_variant("bf")
    md["part"][0]["body"] = np.array([[0], [0], [0]])

    # ---- theta-parameterised solves (main_custom_training.py:32-38) -----------
    M = dg.MeasurementData
    M.theta_mean, M.theta_std = np.array([np.log(20.0), 0.0]), np.array([0.1, 0.015])
    M.node_id, M.ele_id, M.nipt_id = 231, 12, np.array([1, 3], dtype=int)
    xs = [[0.0, 0.0], [1.0, -1.0], [-2.5, 3.0], [0.3, 40.0]]
    rng = np.random.default_rng(0)
    xs += rng.standard_normal((12, 2)).tolist()
    xs = np.asarray(xs, dtype=np.float64)
    ys, hs, us, sigs = [], [], [], []
    for x in xs:
        ys.append(np.asarray(M.fem_f_fun(x), dtype=np.float64).ravel())
        us.append(np.asarray(fp.PreProcessing.sol_data["u_n1"].toarray()).ravel())
        sigs.append(np.asarray(fp.PreProcessing.out_data["ele_stress"][:, :, :, 1]).copy())
        hs.append(np.asarray(M.fem_h_fun(x), dtype=np.float64).ravel())
    out["x"] = xs
    out["y"] = np.stack(ys)
    out["h"] = np.stack(hs)
    out["u"] = np.stack(us)
    out["stress"] = np.stack(sigs)

    # ---- gradient pins: central differences of the reference twin ITSELF (Richardson-extrapolated, two
    #      step sizes) -> the 4x2 Jacobian d(y0, y1, h0, h1)/d(x0, x1) at a few x.  tape.gradient
    #      (main_custom_training.py:252-256) differentiates the TF copy of these formulas; TF cannot run
    #      here, so this is the closest reference-side pin of the gradient (good to ~1e-9 relative).
    def yh(x):
        return np.concatenate([np.asarray(M.fem_f_fun(x), dtype=np.float64).ravel(),
                               np.asarray(M.fem_h_fun(x), dtype=np.float64).ravel()])

    fd_x = np.array([[1.0, -1.0], [0.0, 0.0], [-2.5, 3.0], [0.3, 40.0]])
    fd_jac = np.zeros((len(fd_x), 4, 2))
    for i, x in enumerate(fd_x):
        for k in range(2):
            d = []
            for hstep in (0.02, 0.01):
                xp, xm = x.copy(), x.copy()
                xp[k] += hstep
                xm[k] -= hstep
                d.append((yh(xp) - yh(xm)) / (2.0 * hstep))
            fd_jac[i, :, k] = (4.0 * d[1] - d[0]) / 3.0
    out["fd_x"] = fd_x
    out["fd_jac"] = fd_jac
    np.savez_compressed(os.path.join(HERE, "ref_numpy_twin.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_numpy_twin.npz"))
    print("c1 y", out["c1_nodal_disp"][:, 230], "vm", out["c1_vm"], "tol", out["c1_tol"])
    for x, y, h in zip(xs[:4], ys[:4], hs[:4]):
        print(x, y, h)


if __name__ == "__main__":
    main()
