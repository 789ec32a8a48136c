"""Tiny profiling target: a few fused forward+adjoint launches at the benchmark
size (Cook 20x10, batch 4096), nothing else -- keeps ncu replays short."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import importlib  # noqa: E402

pkg = importlib.import_module(bench.PKG)
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
xh, gyh, ghh = bench.inputs(0)
x, gy, gh = (torch.tensor(a[:n], device=dev) for a in (xh, gyh, ghh))
for _ in range(reps):
    y, h, gx = eng.forward_backward(x, gy, gh)
torch.cuda.synchronize()
bad, _ = eng.status(n)
print("ok", float(y.sum()), float(gx.sum()), "flagged", bad)
sys.exit(1 if bad else 0)
