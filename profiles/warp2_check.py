"""Warp kernel, second generation (K = lambda K_lam + mu K_mu from the band table in shared memory) against the first
(VBFEM_WARP_V1=1: element matrices per sample) on the same inputs: Cook 20x10, max relative differences and
CUDA-event timings.   python profiles/warp2_check.py [n=4096] [reps=20]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
pkg = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g, md = ge._golden_model()
os.environ["VBFEM_WARP_V1"] = "1"
ref = pkg.CookFemEngine(md, device=0)
del os.environ["VBFEM_WARP_V1"]
eng = pkg.CookFemEngine(md, device=0)
print("v1", ref.info["smem_bytes"], "v2", eng.info)
dev = eng.device
rng = np.random.default_rng(4)
x = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gy = torch.tensor(rng.standard_normal((n, 2)), device=dev)
gh = torch.tensor(rng.standard_normal((n, 2)), device=dev)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
y0, h0 = ref.forward(x); y1, h1 = eng.forward(x); torch.cuda.synchronize()
print("fwd  y", rel(y1, y0), "h", rel(h1, h0), "flagged", eng.status(n)[0])
y0, h0, g0 = ref.forward_backward(x, gy, gh); y1, h1, g1 = eng.forward_backward(x, gy, gh); torch.cuda.synchronize()
print("adj  y", rel(y1, y0), "h", rel(h1, h0), "gx", rel(g1, g0), "gx element-wise", float(((g1 - g0).abs() / g0.abs()).max()),
      "flagged", eng.status(n)[0])
_, _, j0 = ref.forward_jac(x); _, _, j1 = eng.forward_jac(x); torch.cuda.synchronize()
print("jac ", rel(j1, j0))
yg = torch.tensor(g["x"], device=dev)
yy, hh = eng.forward(yg); torch.cuda.synchronize()
print("golden y", float(np.max(np.abs(yy.cpu().numpy() - g["y"]) / np.abs(g["y"]))), "h", float(np.max(np.abs(hh.cpu().numpy() - g["h"]) / np.abs(g["h"]))))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
os.environ["VBFEM_WARP2_NW16"] = "0"
e12 = pkg.CookFemEngine(md, device=0)
del os.environ["VBFEM_WARP2_NW16"]
for name, e in (("v1", ref), ("v2 12 warps", e12), ("v2", eng)):
    for mode in ("fwd", "adj", "jac"):
        f = {"fwd": lambda: e.forward(x), "adj": lambda: e.forward_backward(x, gy, gh), "jac": lambda: e.forward_jac(x)}[mode]
        f(); f(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms = tot / reps
        print(f"{name} {mode}: {ms:.4f} ms per launch of {n} = {n / ms / 1e3:.3f} M solves/s")
