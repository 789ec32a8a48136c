import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
PKG = "variational-bayesian-inference-for-computational-mechanics_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def golden():
    """Outputs of the UNMODIFIED reference NumPy twin (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_numpy_twin.npz"))


def model_from_golden(g):
    return {
        "mesh_info": {"nnodes": 231, "nele": 200, "coord": g["coord"]},
        "dof_info": {"IEN": g["IEN"], "LM": g["LM"], "ID": g["ID"], "free_dof": g["free_dof"],
                     "supp_dof": g["supp_dof"], "ndof": 462, "nfree": int(g["free_dof"].size),
                     "nsupp": int(g["supp_dof"].size)},
        "loading": {"Pf": g["Pf"].reshape(-1, 1)},
        "section": [{"thk": 10}],
        "material": [{"E": 20.0, "v": 0.3}],
        "solution_control": {"solver": 1, "load_control": {"numsteps": 1}},
    }


@pytest.fixture(scope="session")
def golden_model(golden):
    return model_from_golden(golden)


@pytest.fixture(scope="session")
def oracle_mesh(golden):
    mesh = {"nnodes": 231, "nele": 200, "coord": golden["coord"], "conn": golden["IEN"]}
    dof = {"IEN": golden["IEN"], "LM": golden["LM"], "free_dof": golden["free_dof"], "ndof": 462,
           "Pf": golden["Pf"]}
    return mesh, dof


@pytest.fixture(scope="session")
def torch_oracle(oracle_mesh):
    import fem_oracle as fo
    return fo.TorchOracle(*oracle_mesh)


@pytest.fixture(scope="session")
def engine(pkg, golden_model):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.CookFemEngine(golden_model, device=0)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
