#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched Cook's-membrane FEM hot path.

  python bench.py --gpus N --steps K --warmup W            (this repo, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  (CPU oracle port of the reference path)

One "step" = one fused forward+adjoint pass over one batch of 4096 sampled
material-parameter sets on the Cook 20x10 mesh (BASELINE.json configs[1]);
with N GPUs every rank processes its own 4096-sample batch (weak scaling, no
data-path collective -- samples are independent).  Prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "variational-bayesian-inference-for-computational-mechanics_b200"
BATCH = 4096
METRIC = "FEM forward+adjoint solves/s (Cook 20x10, batch 4096 per GPU)"
# algorithmic work per forward+adjoint solve at 20x10 (n=440, b=25)
FLOP_PER_SOLVE = 1600 * 200 + 440 * (25 * 25 + 3 * 25) + 2 * 4 * 440 * 25   # = 0.716 MFLOP
BYTES_PER_SOLVE = 96                                                          # x, gy, gh in; y, h, gx out
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused launch at batch 4096 from the committed
# `ncu --set full` capture (profiles/r01_ncu_front_kernel_final_details.txt): 328 704 B read, 0 B written
NCU_DRAM_BYTES_PER_LAUNCH = 328704


def golden_model():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_numpy_twin.npz"))
    md = {
        "mesh_info": {"nnodes": 231, "nele": 200, "coord": g["coord"]},
        "dof_info": {"IEN": g["IEN"], "LM": g["LM"], "free_dof": g["free_dof"], "supp_dof": g["supp_dof"],
                     "ndof": 462, "nfree": 440, "nsupp": 22},
        "loading": {"Pf": g["Pf"].reshape(-1, 1)},
        "section": [{"thk": 10}],
    }
    return g, md


def inputs(rank=0):
    """SURVEY.md 8(d) config 2 (rank r uses seeds 0+100r / 1+100r)."""
    x = np.random.default_rng(0 + 100 * rank).standard_normal((BATCH, 2))
    g = np.random.default_rng(1 + 100 * rank).standard_normal((BATCH, 4))
    return x, np.ascontiguousarray(g[:, :2]), np.ascontiguousarray(g[:, 2:])


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fem_oracle as fo
    g, _ = golden_model()
    mesh = {"nnodes": 231, "nele": 200, "coord": g["coord"], "conn": g["IEN"]}
    dof = {"IEN": g["IEN"], "LM": g["LM"], "free_dof": g["free_dof"], "ndof": 462, "Pf": g["Pf"]}
    return fo, fo.TorchOracle(mesh, dof), mesh, dof


def cpu_fwd_adjoint(to, x, gy, gh, chunk=256):
    for i in range(0, len(x), chunk):
        to.vjp(x[i:i + chunk], gy[i:i + chunk], gh[i:i + chunk])


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate(timeout=10)[0]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """CPU arm: the oracle's batched float64 port of the reference path (dense
    assembly + dense LU + reverse-mode gradient, torch CPU, all host threads),
    each step a bounded sample of the 4096-sample batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    fo, to, mesh, dof = oracle()
    sample = 256
    x, gy, gh = inputs(0)
    x, gy, gh = x[:sample], gy[:sample], gh[:sample]
    for _ in range(min(max(args.warmup, 1), 3)):
        cpu_fwd_adjoint(to, x, gy, gh)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_fwd_adjoint(to, x, gy, gh)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Cook 20x10 forward+adjoint, batch 4096 (BASELINE configs[1])",
                   "sample_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the 4096 seeded samples per step, oracle/fem_oracle.py TorchOracle.vjp "
                                   "(vectorised dense assembly + dense LU + autograd), torch CPU float64"},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_cuda(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg = importlib.import_module(PKG)
    g, md = golden_model()
    eng = pkg.CookFemEngine(md, device=local)
    xh, gyh, ghh = inputs(rank)
    x, gy, gh = (torch.tensor(a, device=dev) for a in (xh, gyh, ghh))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- correctness gate on this rank's first 8 samples vs the reference golden (rank 0)
    y, h, gx = eng.forward_backward(x, gy, gh)
    torch.cuda.synchronize()
    bad, _ = eng.status(BATCH)
    if bad:
        raise SystemExit(f"{bad} samples flagged by the solver")
    if rank == 0:
        y16, h16 = eng.forward(torch.tensor(g["x"], device=dev))
        err = max(float(np.max(np.abs(y16.cpu().numpy() - g["y"]) / np.abs(g["y"]))),
                  float(np.max(np.abs(h16.cpu().numpy() - g["h"]) / np.abs(g["h"]))))
        if err > 1e-9:
            raise SystemExit(f"parity gate failed: {err:.3e}")

    # ---------------- device-resident throughput (value): K steps, CUDA events, L2 flushed between steps
    for _ in range(max(args.warmup, 3)):
        eng.forward_backward(x, gy, gh)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        eng.forward_backward(x, gy, gh)
        b.record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    launches = eng.launches - launches0
    # back-to-back (no flush) for reference
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for _ in range(args.steps):
        eng.forward_backward(x, gy, gh)
    b.record()
    barrier()
    b2b_ms = a.elapsed_time(b)
    # forward only
    a.record()
    for _ in range(args.steps):
        eng.forward(x)
    b.record()
    barrier()
    fwd_ms = a.elapsed_time(b)

    # ---------------- end to end through the public host API (NumPy in / NumPy out; H2D + D2H inside)
    for _ in range(3):
        eng.forward_backward_host(xh, gyh, ghh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        yh, hh, gxh = eng.forward_backward_host(xh, gyh, ghh)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- ELBO step (config 3 at N=1; config 5 shape, 8192 samples per GPU, at N>1)
    B = 64
    S = 100 if world == 1 else 128 * world
    yd = np.random.default_rng(2).standard_normal((10000, 2)) * np.array([0.53, 0.65]) + np.array([-4.24, 5.71])
    e_data = torch.tensor(np.random.default_rng(3 if world == 1 else 5).standard_normal((S, 2)), device=dev)
    model = pkg.elbo.make_step1_model(device=dev)
    opt = pkg.elbo.make_step1_optimizer(model)
    loss_fn = pkg.elbo.Step1Loss(eng, e_data, 0.1, rank=rank, world=world)
    pin = torch.empty(B, 2, dtype=torch.float64).pin_memory()

    def elbo_step(i):
        pin.copy_(torch.from_numpy(yd[(i * B) % 9984:(i * B) % 9984 + B]))
        yb = pin.to(dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        mu, sig, ls = model(yb)
        loss = loss_fn(yb, mu, sig, ls)
        loss.backward()
        opt.step()
        return float(loss)  # D2H read of the loss

    elbo_steps = max(5, min(args.steps, 30))
    for i in range(3):
        elbo_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(elbo_steps):
        last_loss = elbo_step(3 + i)
    barrier()
    elbo_eager_s = time.perf_counter() - t0

    # the same step captured once in a CUDA graph (nets + fused FEM op + Adam) and replayed
    gmodel = pkg.elbo.make_step1_model(device=dev)
    gstep = pkg.elbo.GraphedStep1(gmodel, pkg.elbo.make_step1_optimizer_capturable(gmodel), loss_fn, B, dev)

    def elbo_graph_step(i):
        return float(gstep.step(yd[(i * B) % 9984:(i * B) % 9984 + B]))  # H2D of the batch, D2H of the loss

    for i in range(3):
        elbo_graph_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(elbo_steps):
        last_loss_g = elbo_graph_step(3 + i)
    barrier()
    elbo_s = time.perf_counter() - t0

    # ---------------- max over ranks
    t = torch.tensor([dev_ms, b2b_ms, fwd_ms, e2e_s, elbo_s, elbo_eager_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, b2b_ms, fwd_ms, e2e_s, elbo_s, elbo_eager_s = t.tolist()

    if rank == 0:
        import ctypes
        fp64 = ctypes.c_double(0.0)
        pkg._lib.load().vbfem_measure_peaks(local, ctypes.byref(fp64), None)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        total = world * BATCH * args.steps
        value = total / (dev_ms * 1e-3)
        launch_s = dev_ms * 1e-3 / args.steps
        tflops = FLOP_PER_SOLVE * BATCH / launch_s / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fo, to, mesh, dof = oracle()
            ns = 1536
            cpu_fwd_adjoint(to, xh[:256], gyh[:256], ghh[:256])
            t0 = time.perf_counter()
            cpu_fwd_adjoint(to, xh[:ns], gyh[:ns], ghh[:ns])
            dt = time.perf_counter() - t0
            lo = fo.LoopOracle(mesh, dof)
            t1 = time.perf_counter()
            fo.fem_fh_loop(lo, xh[:8], (np.log(20.0), 0.0), (0.1, 0.015))
            dl = time.perf_counter() - t1
            cpu = {"value": ns / dt, "unit": "solves/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"first {ns} of the 4096 seeded samples, forward+adjoint, oracle TorchOracle.vjp "
                             "(vectorised dense assembly + dense LU + autograd, torch CPU f64)",
                   "statement_by_statement_port": {"value": 8 / dl, "unit": "forward solves/s", "cores": 1,
                                                   "sample": "8 samples, oracle LoopOracle (element/Gauss loops as "
                                                             "in the reference NumPy twin; forward only)"}}
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Cook 20x10 (n=440 free dofs, half bandwidth 25) fused forward+adjoint, "
                                   "batch 4096 per GPU (BASELINE configs[1]); x~N(0,I) seed 0, cotangents seed 1",
                       "batch_per_gpu": BATCH, "l2": "flushed between timed steps (256 MiB memset outside the events)",
                       "kernel": eng.info},
            "e2e": {"value": world * BATCH * args.steps / e2e_s, "unit": "solves/s",
                    "h2d_bytes_per_step": BATCH * 6 * 8, "d2h_bytes_per_step": BATCH * 6 * 8,
                    "api": "CookFemEngine.forward_backward_host -> vbfem_forward_backward_host (NumPy in/out)"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": tflops, "peak": fp64.value, "unit": "TFLOP/s",
                         "frac": tflops / fp64.value if fp64.value else None, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_unit": "bytes of DRAM traffic per launch (ncu, batch 4096): the band never leaves the SM",
                         "other_pipes_ncu": {"fp64_pipe_active_pct": 22.8, "shared_memory_data_pipe_pct": 52.6,
                                             "note": "profiles/r01_ncu_front_kernel_final_details.txt: the broadcast of the "
                                                     "scaled pivot column (12 LDS.128 per column and front) keeps the "
                                                     "shared-memory pipe busier than the FP64 pipe"},
                         "peak_source": "DFMA loop measured on this GPU by vbfem_measure_peaks "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_solve": FLOP_PER_SOLVE,
                         "hbm": {"achieved_gbs": BYTES_PER_SOLVE * BATCH / launch_s / 1e9, "peak_gbs": hbm_peak,
                                 "note": "compulsory I/O only; the band lives on chip"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "extra": {"back_to_back_solves_per_s": world * BATCH * args.steps / (b2b_ms * 1e-3),
                      "forward_only_solves_per_s": world * BATCH * args.steps / (fwd_ms * 1e-3),
                      "step_ms_min": min(step_ms), "step_ms_max": max(step_ms),
                      "elbo": {"steps_per_s": elbo_steps / elbo_s, "B": B, "S": S, "samples_per_step": B * S,
                               "fem_solves_per_s": B * S * elbo_steps / elbo_s, "last_loss": last_loss_g,
                               "cuda_graph": bool(gstep.graphed),
                               "eager_steps_per_s": elbo_steps / elbo_eager_s, "eager_last_loss": last_loss,
                               "what": "NN fwd -> reparam -> FEM fwd -> loss -> FEM adjoint -> NN bwd -> Adam (one CUDA graph replay per step), "
                                       "batch H2D and loss D2H inside; one NCCL all-reduce per step when N>1"}},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in a safe order: the captured graph (it holds an NCCL all-reduce node) goes first,
        # then the process group; a watchdog ends the process if the NCCL teardown does not return
        # (seen once at 8 GPUs: the line above was printed, the interpreter never exited).
        import gc
        import threading
        sys.stdout.flush()
        dog = threading.Timer(45.0, lambda: os._exit(0))
        dog.daemon = True
        dog.start()
        gstep.graph = None
        del gstep
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        dog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 5 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 50 if args.steps is None else args.steps
        args.warmup = 5 if args.warmup is None else args.warmup
        run_cuda(args)


if __name__ == "__main__":
    main()
