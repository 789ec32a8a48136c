"""Config 4: Cook 80x40, batch 1024 forward+adjoint, x ~ default_rng(4)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")
P = pkg.PreProcessing
md = P.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
print(eng.info)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = eng.device
x = torch.tensor(np.random.default_rng(4).standard_normal((n, 2)), device=dev)
gy = torch.ones(n, 2, dtype=torch.float64, device=dev); gh = torch.full((n, 2), 0.5, dtype=torch.float64, device=dev)
for mode in ("fwd", "fwd+adj"):
    f = (lambda: eng.forward(x)) if mode == "fwd" else (lambda: eng.forward_backward(x, gy, gh))
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nfree, b = eng.info["nfree"], eng.info["half_bw"]
    flop = 1600 * eng.info["nele"] + nfree * (b * b + 3 * b) + (2 if mode != "fwd" else 1) * 4 * nfree * b
    print(f"{mode}: {n} samples in {dt*1e3:.1f} ms = {n/dt:.0f} solves/s, {flop*n/dt/1e12:.3f} TFLOP/s algorithmic, bad={eng.status(n)[0]}")
