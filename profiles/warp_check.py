"""Warp-per-sample kernel against the on-chip two-front kernel (Cook 20x10): same inputs, both engines in one
process (VBFEM_WARP toggled around engine creation), max relative differences and CUDA-event timings.
  python profiles/warp_check.py [n=4096] [reps=10]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
g, md = bench.golden_model()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
os.environ["VBFEM_WARP"] = "0"
ref = pkg.CookFemEngine(md, device=0)
os.environ["VBFEM_WARP"] = "1"
eng = pkg.CookFemEngine(md, device=0)
print("ref", ref.info["kernel_variant"], "new", eng.info)
dev = eng.device
rng = np.random.default_rng(0)
x = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gy = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gh = torch.tensor(rng.standard_normal((n, 2)), device=dev)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
y0, h0 = ref.forward(x)
y1, h1 = eng.forward(x)
torch.cuda.synchronize()
print("fwd  y", rel(y1, y0), "h", rel(h1, h0), "flagged", eng.status(n)[0])
y0, h0, g0 = ref.forward_backward(x, gy, gh)
y1, h1, g1 = eng.forward_backward(x, gy, gh)
torch.cuda.synchronize()
print("adj  y", rel(y1, y0), "h", rel(h1, h0), "gx", rel(g1, g0), "gx1", rel(g1[:, 1], g0[:, 1]), "flagged", eng.status(n)[0])
_, _, j0 = ref.forward_jac(x)
_, _, j1 = eng.forward_jac(x)
torch.cuda.synchronize()
print("jac ", rel(j1, j0))
for name, e in (("front", ref), ("warp", eng)):
    for mode in ("fwd", "adj", "jac"):
        f = {"fwd": lambda: e.forward(x), "adj": lambda: e.forward_backward(x, gy, gh), "jac": lambda: e.forward_jac(x)}[mode]
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{name:6s} {mode}: {ms:.3f} ms per launch of {n} = {n / ms / 1e3:.3f} M solves/s")
