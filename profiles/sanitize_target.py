"""compute-sanitizer target: small launches of every production kernel.

  python profiles/sanitize_target.py front   [n]  -- Cook 20x10: forward, fused forward+adjoint, Jacobian mode + backward
  python profiles/sanitize_target.py big     [n]  -- Cook 80x40 (band in HBM): forward, fused forward+adjoint
  python profiles/sanitize_target.py fields       -- full-field mode of fem_test.py (generic kernel, band in shared memory)

Used as `compute-sanitizer --tool racecheck|memcheck python profiles/sanitize_target.py ...`;
the logs are committed under profiles/.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module(bench.PKG)
what = sys.argv[1] if len(sys.argv) > 1 else "front"
dev = torch.device("cuda", 0)
t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)

if what in ("front", "fields"):
    g, md = bench.golden_model()
    eng = pkg.CookFemEngine(md, device=0)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 296
else:
    md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
    eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
print(eng.info)
x = t(np.random.default_rng(0).standard_normal((n, 2)))
gy = t(np.random.default_rng(1).standard_normal((n, 2)))
gh = t(np.random.default_rng(2).standard_normal((n, 2)))
if what == "fields":
    out = eng.fields(x=x[:2])
    torch.cuda.synchronize()
    print("fields ok", float(out["u"].abs().sum()))
else:
    y, h = eng.forward(x)
    y1, h1, gx = eng.forward_backward(x, gy, gh)
    y2, h2 = eng.forward(x, keep_factor=True)
    gx2 = eng.backward(gy, gh)
    torch.cuda.synchronize()
    bad, _ = eng.status(n)
    err = float((gx - gx2).abs().max() / gx.abs().max())
    print("ok", float(y.sum()), float(gx.sum()), "fused-vs-jacobian", err, "flagged", bad)
    sys.exit(1 if bad or not err < 1e-9 else 0)
