"""Phase budget of the blocked panel kernel (Cook 80x40) from clock64 accumulators (profiling build
libvbfem_tl.so, -DVBFEM_TIMELINE, selected with VBFEM_LIB): per warp role, the SM cycles the LAST
sample of every CTA spent in each phase, work and barrier wait separately."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
md = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
eng = pkg.CookFemEngine(md, device=0, node_id=3321, ele_id=12)
dev = eng.device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mode = sys.argv[2] if len(sys.argv) > 2 else "adj"
x = torch.tensor(np.random.default_rng(4).standard_normal((n, 2)), device=dev)
gy = torch.ones(n, 2, dtype=torch.float64, device=dev)
gh = torch.full((n, 2), 0.5, dtype=torch.float64, device=dev)
for _ in range(2):
    eng.forward_backward(x, gy, gh) if mode == "adj" else eng.forward(x)
torch.cuda.synchronize()
lib = pkg._lib.load()
ncta = eng.info["num_sms"] * eng.info["ctas_per_sm"]
buf = np.zeros(ncta * 4 * 16, dtype=np.int64)
lib.vbfem_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
lib.vbfem_debug_timeline(eng._h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(ncta, 4, 16).astype(np.float64)[:min(n, ncta)]
names = ["reset + first window rows + first diagonal block", "B work (panel solve)", "B barrier", "C work (update | look-ahead | assembly)",
         "C barrier", "-", "-", "observations + reverse-pass set-up", "reverse pass", "contraction + store",
         "  asm: clear + rhs block", "  asm: issue table loads", "  asm: element matrices", "  asm: gather", "  asm: rotate prefetch registers", "-"]
roles = ["warp 0 (solve + update)", "warp 1 (solve + update)", "warp 6 (look-ahead diagonal block, bulk store)", "warp 7 (assembly)"]
npan = 820
for w in range(4):
    print(f"--- {roles[w]}   (mean over {t.shape[0]} CTAs; per panel = / {npan})")
    tot = t[:, w, :13].sum(axis=1).mean() if w == 2 else t[:, w, :10].sum(axis=1).mean()
    names_w = names if w == 3 else (names[:10] + ["  look-ahead: bulk store + wait for the other staging buffer", "  look-ahead: update of the next diagonal block", "  look-ahead: LDL^T + inverse of the next diagonal block"] if w == 2 else names[:10])
    for i, nm in enumerate(names_w):
        v = t[:, w, i].mean()
        print(f"  {nm:48s} {v:12.0f} cycles  {100 * v / tot:5.1f} %   per panel {v / npan:8.1f}")
    print(f"  {'total':48s} {tot:12.0f}")
