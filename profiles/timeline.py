"""Phase timeline of the front kernel from clock64 marks (profiling build libvbfem_tl.so,
compiled with -DVBFEM_TIMELINE; select it with VBFEM_LIB).  Prints, per warp role, the mean
duration in SM cycles of every phase of the LAST sample each CTA processed."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
dev = torch.device("cuda", 0)
xh, gyh, ghh = bench.inputs(0)
x, gy, gh = (torch.tensor(a, device=dev) for a in (xh, gyh, ghh))
for _ in range(3):
    eng.forward_backward(x, gy, gh)
torch.cuda.synchronize()
lib = pkg._lib.load()
ncta = eng.info["num_sms"] * eng.info["ctas_per_sm"]
buf = np.zeros(ncta * 4 * 16, dtype=np.int64)
lib.vbfem_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
rc = lib.vbfem_debug_timeline(eng._h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(ncta, 4, 16).astype(np.float64)
names = {0: "start", 1: "zeroed", 2: "assembled", 3: "own columns eliminated", 4: "hand-over barrier passed",
         5: "middle eliminated", 6: "u_mid, y, h", 7: "psi_mid / D^-1 done", 8: "release barrier passed",
         9: "joint back substitution", 10: "front phase end (block barrier)", 11: "contraction + store"}
# roles: a front warp has mark 3; the top front has mark 5
role = np.where(t[:, :, 5] > 0, 0, np.where(t[:, :, 3] > 0, 1, 2))
for r, nm in ((0, "TOP front (finishes the middle)"), (1, "BOTTOM front (unit vectors)"), (2, "helper warps")):
    sel = role == r
    print(f"--- {nm}: {int(sel.sum())} warps")
    prev = t[:, :, 0]
    for i in range(1, 12):
        cur = t[:, :, i]
        ok = sel & (cur > 0) & (prev > 0)
        if ok.any():
            d = (cur - prev)[ok]
            print(f"  {names[i]:34s} {d.mean():9.0f} cycles  (min {d.min():7.0f} max {d.max():7.0f})  cumulative {(cur - t[:, :, 0])[ok].mean():9.0f}")
            prev = np.where(cur > 0, cur, prev)
