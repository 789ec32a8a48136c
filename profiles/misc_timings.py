"""Timings of the secondary paths DESIGN.md quotes: the generic kernel (full fields, fem_test.py's path; its
band-in-shared-memory instantiation spills), Jacobian mode, the front kernel, forced-generic forward/adjoint.
CUDA events, median of 5 launches after a warm-up.   python profiles/misc_timings.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
g, md = bench.golden_model()


def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


n = 4096
xh, gyh, ghh = bench.inputs(0)
for label, env in (("warp kernel (default)", {}), ("front kernel (VBFEM_WARP=0)", {"VBFEM_WARP": "0"}),
                   ("panel kernel at 20x10 (VBFEM_FORCE_PANEL=1)", {"VBFEM_FORCE_PANEL": "1"}),
                   ("generic kernel (VBFEM_FORCE_GENERIC=1)", {"VBFEM_FORCE_GENERIC": "1"})):
    for k, v in env.items():
        os.environ[k] = v
    eng = pkg.CookFemEngine(md, device=0)
    for k in env:
        del os.environ[k]
    x, gy, gh = (torch.tensor(a, device=eng.device) for a in (xh, gyh, ghh))
    row = [f"{label:46s} variant {eng.info['kernel_variant']}"]
    for mode, f in (("fwd", lambda: eng.forward(x)), ("fwd+adj", lambda: eng.forward_backward(x, gy, gh)),
                    ("jacobian", lambda: eng.forward_jac(x))):
        ms = timeit(f)
        row.append(f"{mode} {n / ms / 1e3:7.3f} M/s")
    print("  ".join(row))
    if not env:
        emat = torch.tensor(np.tile([[20.0, 0.3]], (n, 1)), device=eng.device)
        for want in (("u",), ("u", "stress", "strain", "fint")):
            ms = timeit(lambda: eng.fields(emat=emat, want=want))
            print(f"  vbfem_fields (generic kernel, band in shared memory), {n} samples, outputs {want}: {ms:.2f} ms = "
                  f"{n / ms / 1e3:.3f} M solves/s")
        ms = timeit(lambda: eng.fields(emat=emat[:1], want=("u", "stress", "strain", "fint")), reps=20)
        print(f"  vbfem_fields, 1 sample (fem_test.py): {ms * 1e3:.0f} us")
    eng.close()
md4 = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
eng = pkg.CookFemEngine(md4, device=0, node_id=3321, ele_id=12)
x = torch.tensor(np.random.default_rng(4).standard_normal((1024, 2)), device=eng.device)
print(f"80x40 panel kernel, Jacobian mode, 1024 samples: {1024 / timeit(lambda: eng.forward_jac(x)):.1f} k solves/s")
emat = torch.tensor(np.tile([[20.0, 0.3]], (64, 1)), device=eng.device)
ms = timeit(lambda: eng.fields(emat=emat, want=("u",)))
print(f"80x40 vbfem_fields (generic kernel, band in HBM), 64 samples: {ms:.1f} ms = {64 / ms:.2f} k solves/s")
