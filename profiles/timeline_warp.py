"""Phase budget of the warp-per-sample kernel (Cook 20x10) from clock64 accumulators (profiling build
libvbfem_tl.so, -DVBFEM_TIMELINE, selected with VBFEM_LIB): SM cycles the LAST sample of warps 0..3 of every
CTA spent in each phase (all 12 warps of the SM running)."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
os.environ["VBFEM_WARP"] = "1"
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
dev = eng.device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "adj"
x = torch.tensor(np.random.default_rng(0).standard_normal((n, 2)), device=dev)
gy = torch.ones(n, 2, dtype=torch.float64, device=dev)
gh = torch.full((n, 2), 0.5, dtype=torch.float64, device=dev)
for _ in range(2):
    eng.forward_backward(x, gy, gh) if mode == "adj" else eng.forward(x)
torch.cuda.synchronize()
lib = pkg._lib.load()
ncta = eng.info["num_sms"]
buf = np.zeros(ncta * 4 * 16, dtype=np.int64)
lib.vbfem_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
lib.vbfem_debug_timeline(eng._h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(ncta, 4, 16).astype(np.float64).reshape(-1, 16)
t = t[t.sum(1) > 0]
names = ["window fill (first 4 rows: elements, gather) + first diagonal block", "loop top / observations",
         "next diagonal block (LDL^T, inverse) || gather of the entering row (+ element batch)",
         "solve + G + panel store", "trailing update", "reverse-pass set-up", "reverse pass", "contraction + store",
         "entering row: fragments from the staging area"]
idx = [0, 1, 2, 3, 4, 8, 5, 6, 7]
if eng.info["panel_ring"] == 0:   # second generation (vbfem_warp2.cuh)
    names = ["window fill (first 4 rows from the band table) + first diagonal block", "loop top / observations",
             "entering row from the band table || next diagonal block (LDL^T, inverse)", "solve + G + panel store",
             "trailing update", "reverse-pass set-up", "final reduction + store", "status", "-",
             "reverse pass: slab loads + back substitution MMAs", "reverse pass: window write + band contraction"]
    idx = [0, 1, 2, 3, 4, 5, 9, 10, 6, 7]
tot = t[:, :11].sum(1).mean()
print(f"{mode}: {t.shape[0]} warps, mean cycles per sample {tot:.0f}  ({eng.info})")
for i in idx:
    v = t[:, i].mean()
    print(f"  {names[i]:84s} {v:10.0f} cycles {100 * v / tot:5.1f} %   per panel {v / 55:8.1f}")
