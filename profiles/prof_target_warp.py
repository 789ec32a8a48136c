"""Profiling target: Cook 20x10, the warp-per-sample kernel, a few fused forward+adjoint launches at batch 4096.
  python profiles/prof_target_warp.py [mode=adj|fwd|jac]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
os.environ.setdefault("VBFEM_WARP", "1")
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
mode = sys.argv[1] if len(sys.argv) > 1 else "adj"
x, gy, gh = (torch.tensor(a, device=eng.device) for a in bench.inputs(0))
for _ in range(2):
    out = eng.forward_backward(x, gy, gh) if mode == "adj" else (eng.forward_jac(x) if mode == "jac" else eng.forward(x))
torch.cuda.synchronize()
print("ok", mode, eng.info["kernel_variant"], float(out[0].sum()), "flagged", eng.status(4096)[0])
