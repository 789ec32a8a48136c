"""Launch time of the warp-per-sample kernel against the batch size (quantisation of the 148 x NW warp slots).
  python profiles/warp_sweep.py mode n1 n2 ..."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
os.environ.setdefault("VBFEM_WARP", "1")
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
mode = sys.argv[1]
print(eng.info)
for n in [int(v) for v in sys.argv[2:]]:
    rng = np.random.default_rng(0)
    x, gy, gh = (torch.tensor(rng.standard_normal((n, 2)), device=eng.device) for _ in range(3))
    f = (lambda: eng.forward_backward(x, gy, gh)) if mode == "adj" else (lambda: eng.forward(x))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        f()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{mode} n={n:6d}: {ms:.3f} ms = {n / ms / 1e3:.3f} M solves/s  ({n / 148 / (eng.info['block_threads'] // 32):.2f} rounds)")
