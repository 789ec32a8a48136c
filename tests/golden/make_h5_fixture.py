"""Recipe of tests/golden/chunked_filters.h5: a small file in the layout of the reference's shipped
data_fem_test_big_noise.h5 (MATLAB-7.3 user block, version-0 superblock, chunked float64 datasets stored
transposed, shuffle + deflate + fletcher32), written by h5io.write(compress=True) from seeded arrays.
When /root/reference is present the script also checks that h5io.read decodes the shipped file to the
statistics SURVEY.md section 4 records (y = -4.23 +- 0.53 / 5.71 +- 0.65, 10000 x 2).
Usage: python tests/golden/make_h5_fixture.py"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
h5io = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200.h5io")

rng = np.random.default_rng(7)
d = {"a_data": rng.standard_normal((300, 2)), "b_mean": rng.standard_normal((1, 2))}
out = os.path.join(HERE, "chunked_filters.h5")
h5io.write(data=d, filename=out, compress=True)
back = h5io.read(filename=out)
assert all(np.array_equal(back[k], d[k]) for k in d)
print("wrote", out, os.path.getsize(out), "bytes")
ref = "/root/reference/data_fem_test_big_noise.h5"
if os.path.exists(ref):
    r = h5io.read(filename=ref)
    assert r["y_data"].shape == (10000, 2) and len(r) == 10
    assert np.allclose(r["y_data"].mean(0), [-4.2314, 5.7139], atol=1e-3) and np.allclose(r["y_mean"][0], r["y_data"].mean(0))
    assert np.allclose(r["log_z_data"], np.log(r["z_data"]))
    print("reference file decoded:", {k: v.shape for k, v in r.items()})
