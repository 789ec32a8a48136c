"""Config 4: Cook 80x40, batch 1024 forward+adjoint, x ~ default_rng(4).  CUDA-event timing,
median of `reps` launches after a warm-up."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")
P = pkg.PreProcessing
md = P.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
print(eng.info)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = eng.device
x = torch.tensor(np.random.default_rng(4).standard_normal((n, 2)), device=dev)
gy = torch.ones(n, 2, dtype=torch.float64, device=dev); gh = torch.full((n, 2), 0.5, dtype=torch.float64, device=dev)
for mode in ("fwd", "fwd+adj", "jac"):
    f = {"fwd": lambda: eng.forward(x), "fwd+adj": lambda: eng.forward_backward(x, gy, gh), "jac": lambda: eng.forward_jac(x)}[mode]
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    dt = float(np.median(ts))
    nfree, b_ = eng.info["nfree"], eng.info["half_bw"]
    flop = 1600 * eng.info["nele"] + nfree * (b_ * b_ + 3 * b_) + (2 if mode != "fwd" else 1) * 4 * nfree * b_
    print(f"{mode}: {n} samples in {dt*1e3:.2f} ms (min {min(ts)*1e3:.2f}) = {n/dt:.0f} solves/s, {flop*n/dt/1e12:.3f} TFLOP/s algorithmic, bad={eng.status(n)[0]}")
