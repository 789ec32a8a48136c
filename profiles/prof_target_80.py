"""Profiling target: Cook 80x40 (config 4), a few launches of one mode, nothing else.
  python profiles/prof_target_80.py [n=296] [mode=adj|fwd|jac] [reps=2]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
pkg = importlib.import_module(bench.PKG)
md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mode = sys.argv[2] if len(sys.argv) > 2 else "adj"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = eng.device
x = torch.tensor(np.random.default_rng(4).standard_normal((n, 2)), device=dev)
gy = torch.tensor(np.random.default_rng(5).standard_normal((n, 2)), device=dev)
gh = torch.tensor(np.random.default_rng(6).standard_normal((n, 2)), device=dev)
for _ in range(reps):
    if mode == "adj":
        out = eng.forward_backward(x, gy, gh)
    elif mode == "jac":
        out = eng.forward_jac(x)
    else:
        out = eng.forward(x)
torch.cuda.synchronize()
bad, _ = eng.status(n)
print("ok", mode, n, float(out[0].sum()), "flagged", bad)
sys.exit(1 if bad else 0)
