"""ELBO training step (B = 64, S = 128: 8192 samples) on one GPU: steps/s of the captured step with the two nets on
parallel graph branches / in sequence and with the fused / foreach Adam.   python profiles/elbo_time.py [steps=60]"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
g, md = bench.golden_model()
eng = pkg.CookFemEngine(md, device=0)
dev = eng.device
B, S = 64, 128
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
yd = np.random.default_rng(2).standard_normal((10000, 2)) * np.array([0.53, 0.65]) + np.array([-4.24, 5.71])
e_data = torch.tensor(np.random.default_rng(5).standard_normal((S, 2)), device=dev)
for par, fused, floss in ((True, True, True), (True, True, False), (True, False, False), (False, True, False),
                          (False, False, False)):
    if True:
        model = pkg.elbo.make_step1_model(device=dev)
        model.parallel_nets = par
        opt = pkg.elbo.make_step1_optimizer_capturable(model) if fused else torch.optim.Adam(
            model.parameters(), lr=1e-3, betas=(0.99, 0.999), eps=1e-10, capturable=True)
        step = pkg.elbo.GraphedStep1(model, opt, pkg.elbo.Step1Loss(eng, e_data, 0.1, fused=floss), B, dev)
        for i in range(3):
            float(step.step(yd[i * B:(i + 1) * B]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tk = [step.step_async(yd[((3 + i) * B) % 9984:((3 + i) * B) % 9984 + B]) for i in range(n)]
        last = step.loss_of(tk[-1])
        dt = time.perf_counter() - t0
        print(f"loss in library {floss!s:5s} parallel nets {par!s:5s} fused Adam {fused!s:5s} graphed {step.graphed}: {n / dt:8.1f} steps/s  "
              f"({1e3 * dt / n:.4f} ms per step)  loss {last:.12f}  {getattr(step, 'capture_error', '')}")
