"""Panel kernel, second generation (window in registers) against the first (VBFEM_PANEL_V1=1) on the same inputs:
Cook 80x40 (and a smaller wide-band mesh), max relative differences and CUDA-event timings.
  python profiles/panel_check.py [n=1024] [reps=5] [nx=80] [ny=40]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 80
ny = int(sys.argv[4]) if len(sys.argv) > 4 else 40
md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(nx, ny))
kw = dict(device=0, node_id=(nx + 1) * (ny + 1), ele_id=12)
os.environ["VBFEM_PANEL_V1"] = "1"
ref = pkg.CookFemEngine(md, **kw)
del os.environ["VBFEM_PANEL_V1"]
eng = pkg.CookFemEngine(md, **kw)
print("ref", ref.info["kernel_variant"], ref.info["smem_bytes"], "new", eng.info)
dev = eng.device
rng = np.random.default_rng(4)
x = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gy = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gh = torch.tensor(rng.standard_normal((n, 2)), device=dev)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
y0, h0 = ref.forward(x); y1, h1 = eng.forward(x); torch.cuda.synchronize()
print("fwd  y", rel(y1, y0), "h", rel(h1, h0), "flagged", eng.status(n)[0])
y0, h0, g0 = ref.forward_backward(x, gy, gh); y1, h1, g1 = eng.forward_backward(x, gy, gh); torch.cuda.synchronize()
print("adj  y", rel(y1, y0), "h", rel(h1, h0), "gx", rel(g1, g0), "flagged", eng.status(n)[0])
_, _, j0 = ref.forward_jac(x); _, _, j1 = eng.forward_jac(x); torch.cuda.synchronize()
print("jac ", rel(j1, j0))
for name, e in (("v1", ref), ("v2", eng)):
    for mode in ("fwd", "adj", "jac"):
        f = {"fwd": lambda: e.forward(x), "adj": lambda: e.forward_backward(x, gy, gh), "jac": lambda: e.forward_jac(x)}[mode]
        f(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{name} {mode}: {ms:.2f} ms per launch of {n} = {n / ms:.1f} k solves/s")
