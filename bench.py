#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched Cook's-membrane FEM hot path.

  python bench.py --gpus N --steps K --warmup W            (this repo, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  (CPU oracle port of the reference path)

One "step" = one fused forward+adjoint pass over one batch of 4096 sampled
material-parameter sets on the Cook 20x10 mesh (BASELINE.json configs[1]);
with N GPUs every rank processes its own 4096-sample batch (weak scaling, no
data-path collective -- samples are independent).  The same JSON line carries the
second half of BASELINE.json's metric as a first-class object, "elbo": ELBO
training steps/s at 8192 Monte-Carlo samples per GPU (B = 64, S = 128 N), whose
timed region contains the NCCL all-reduce of the variational-parameter
gradients; "config4": the refined 80x40 mesh at batch 1024; "latency": the
one-sample-at-a-time host path; and the CPU baselines.  Prints ONE JSON line.
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
# `ncu --set full` capture of the warp-per-sample kernel (profiles/r02_ncu_warp_adj_raw_selected.txt):
# 371.7 MB read + 534.7 MB written -- the factor slab of the adjoint (141 KB per sample, written once, read once);
# the forward-only launch moves the compulsory 48 B per sample.
NCU_DRAM_BYTES_PER_LAUNCH = 411526400 + 542349824   # read + written, profiles/r02_ncu_warp_adj_raw_selected.txt
# config 4 (80x40): DRAM bytes per fused forward+adjoint SOLVE from the committed capture of the panel kernel
NCU_C4_DRAM_BYTES_PER_SOLVE = int((1.871130e9 + 1.928987e9) / 296)   # 12.84 MB (read 6.32 + written 6.52)


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


# ------------------------------------------------------------------------------------------------
# CPU baselines beside the GPU numbers
# ------------------------------------------------------------------------------------------------
def host_cores():
    """Host cores this process may use: affinity mask, capped by the cgroup CPU quota (the GPU boxes expose
    every core of the machine to os.cpu_count() but give the container a slice) and by 32."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 32))


def tf_status():
    """the reference TF path, if TensorFlow exists on the box."""
    try:
        import tensorflow  # noqa: F401
    except Exception as exc:  # noqa: BLE001
        return f"unavailable ({type(exc).__name__}: the reference TF path cannot be timed on this box)"
    return "importable (not timed: the reference tree is not on this box)" if not os.path.isdir("/root/reference") \
        else "importable"


def _twin_worker(xs):
    """One process: the UNMODIFIED reference NumPy twin (src/fem_solver.py + src/mat_subroutine.py) behind
    the import shims of tests/golden/make_golden.py, one forward solve pair per sample."""
    import contextlib
    import io
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg
    mg._shim()
    work = tempfile.mkdtemp(prefix="vbfem_twin_")
    shutil.copy(os.path.join(mg.REF, "Armero_cooksm_20x10.txt"), work)
    os.chdir(work)
    with contextlib.redirect_stdout(io.StringIO()):
        import fem_preprocess as fp
        from src import data_generation_2sam_more_loss as dg
        fp.PreProcessing.modeldata_initialization_topopt("Armero_cooksm_20x10.txt", "model_file.mat")
        M = dg.MeasurementData
        M.theta_mean, M.theta_std = np.array([np.log(20.0), 0.0]), np.array([0.1, 0.015])
        M.node_id, M.ele_id, M.nipt_id = 231, 12, np.array([1, 3], dtype=int)
        t0 = time.perf_counter()
        for x in xs:
            M.fem_f_fun(np.asarray(x))   # one solve (2 assemblies + spsolve): y; h comes from the same state
        return time.perf_counter() - t0


def time_numpy_twin(per_core=8):
    """the reference's own NumPy twin on all host cores (one process per core).  The
    reference tree does not travel to the GPU box; there this reports its absence explicitly."""
    if not os.path.isdir("/root/reference"):
        return "reference tree absent on this box (the statement-by-statement port stands in: " \
               "cpu_baseline.statement_by_statement_port)"
    import multiprocessing as mp
    cores = host_cores()
    x = np.random.default_rng(0).standard_normal((cores * per_core, 2))
    t0 = time.perf_counter()
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"   # one core per worker process
    with mp.get_context("spawn").Pool(cores) as pool:
        inner = pool.map(_twin_worker, [x[i::cores] for i in range(cores)])
    wall = time.perf_counter() - t0
    return {"value": len(x) / max(inner), "unit": "forward solves/s", "cores": cores, "kind": "reference",
            "sample": f"{len(x)} seeded samples, MeasurementData.fem_f_fun of the unmodified reference "
                      f"(src/fem_solver.py NumPy twin), one process per core; wall incl. start-up {wall:.1f} s"}


def _loop_worker(xs):
    fo, _, mesh, dof = oracle()
    lo = fo.LoopOracle(mesh, dof)
    t0 = time.perf_counter()
    fo.fem_fh_loop(lo, np.asarray(xs), (np.log(20.0), 0.0), (0.1, 0.015))
    return time.perf_counter() - t0


def time_loop_port(per_core=4):
    """The statement-by-statement port of the reference NumPy twin (oracle LoopOracle: element and Gauss
    loops as upstream) on all host cores, one process per core."""
    import multiprocessing as mp
    cores = host_cores()
    x = np.random.default_rng(0).standard_normal((cores * per_core, 2))
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"   # one core per worker process
    with mp.get_context("spawn").Pool(cores) as pool:
        inner = pool.map(_loop_worker, [x[i::cores] for i in range(cores)])
    return {"value": len(x) / max(inner), "unit": "forward solves/s", "cores": cores, "kind": "port",
            "sample": f"{len(x)} seeded samples, oracle LoopOracle (element/Gauss loops as in the reference NumPy "
                      "twin; forward only), one process per core"}


def cpu_elbo_step(to, fo, B=64, S=4, steps=3):
    """One ELBO training step on the CPU port: nets -> reparameterisation -> oracle FEM (dense LU) -> loss ->
    autograd -> Adam, on a bounded B*S."""
    import torch
    pkg = importlib.import_module(PKG)
    model = pkg.elbo.make_step1_model()
    opt = pkg.elbo.make_step1_optimizer(model)
    yd = torch.tensor(np.random.default_rng(2).standard_normal((B, 2)) * np.array([0.53, 0.65])
                      + np.array([-4.24, 5.71]))
    e = torch.tensor(np.random.default_rng(5).standard_normal((S, 2)))

    def step():
        opt.zero_grad(set_to_none=True)
        mu, sig, ls = model(yd)
        loss, *_ = fo.elbo_step1_torch(to, yd, mu, sig, e, 0.1)
        loss.backward()
        opt.step()
        return float(loss.detach())
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"solves_per_s": B * S / dt, "B": B, "S": S, "s_per_step": dt, "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"B={B}, S={S} ({B * S} fwd+adjoint solves per step), oracle "
                                      "elbo_step1_torch + torch autograd + Adam, torch CPU f64"}


def run_reference(args):
    """CPU arm: the oracle's batched float64 port of the reference path (dense
    assembly + dense LU + reverse-mode gradient, torch CPU, all host threads),
    each step a bounded sample of the 4096-sample batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core at every N.  The thread count must be
    # fixed through the environment BEFORE torch is imported (torch.set_num_threads breaks MKL's batched LU here:
    # "Parameter 6 was incorrect on entry to DLASWP").
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(host_cores())
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
        "numpy_twin": time_numpy_twin(16),
        "tf": tf_status(),
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

    # ---------------- sustained figure: back-to-back fused launches for at least one second
    n_sus = max(50, int(1.1e3 / max(b2b_ms / args.steps, 1e-3)))
    barrier()
    a.record()
    for _ in range(n_sus):
        eng.forward_backward(x, gy, gh)
    b.record()
    barrier()
    sus_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- latency path: the one-sample-at-a-time callers (src/postprocess_lib.py:78-103)
    lat = {}
    for nb in (1, 8, 64):
        xs = xh[:nb]
        for _ in range(20):
            eng.forward_host(xs)
        reps = 300
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.forward_host(xs)
        lat[f"forward_host_n{nb}_us_per_call"] = 1e6 * (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.forward_backward_host(xs, gyh[:nb], ghh[:nb])
        lat[f"forward_backward_host_n{nb}_us_per_call"] = 1e6 * (time.perf_counter() - t0) / reps
    lat["what"] = ("NumPy in / NumPy out through vbfem_forward_host / vbfem_forward_backward_host, wall clock per call "
                   "incl. launch + synchronise; batches <= 64 run on mapped pinned memory (no staging copies)")

    # ---------------- ELBO step: 8192 Monte-Carlo samples per GPU at every N (B = 64, S = 128 N; config 5 at
    #                  N = 8), the NCCL all-reduce of the variational-parameter gradients inside the timed region
    B = 64
    yd = np.random.default_rng(2).standard_normal((10000, 2)) * np.array([0.53, 0.65]) + np.array([-4.24, 5.71])

    # the ranks' exchange: P2P stores into the peers' mailboxes from inside the reduction kernel (NVLink); NCCL if the
    # box does not allow CUDA IPC between the rank processes
    peer_note = None
    if world > 1:
        os.environ.setdefault("VBFEM_PEER_TIMEOUT_MS", "3000")
        try:
            eng.peer_connect_group(cap_doubles=3 + 4 * B)
            # self-test against NCCL before the exchange is trusted with the training step
            probe = torch.tensor(np.random.default_rng(50 + rank).standard_normal(3 + 4 * B), device=dev)
            want = probe.clone()
            eng.peer_allreduce(probe)
            dist.all_reduce(want)
            eng.peer_status()
            ok = torch.tensor([float((probe - want).abs().max() <= 1e-13 * want.abs().max())], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) != 1.0:
                raise RuntimeError("peer all-reduce disagrees with NCCL")
        except Exception as exc:   # noqa: BLE001 -- reported in the JSON line
            peer_note = repr(exc)[:200]
            eng.peer_world = 0
    use_peer = world > 1 and eng.peer_world == world

    def time_elbo(S, seed, graphed=True, pipelined=True, peer=None):
        e_data = torch.tensor(np.random.default_rng(seed).standard_normal((S, 2)), device=dev)
        loss_fn = pkg.elbo.Step1Loss(eng, e_data, 0.1, rank=rank, world=world, peer=peer)
        model = pkg.elbo.make_step1_model(device=dev)
        if graphed:
            step = pkg.elbo.GraphedStep1(model, pkg.elbo.make_step1_optimizer_capturable(model), loss_fn, B, dev)
            run = lambda i: float(step.step(yd[(i * B) % 9984:(i * B) % 9984 + B]))  # batch H2D, loss D2H
        else:
            opt = pkg.elbo.make_step1_optimizer(model)
            pin = torch.empty(B, 2, dtype=torch.float64).pin_memory()

            def run(i):
                pin.copy_(torch.from_numpy(yd[(i * B) % 9984:(i * B) % 9984 + B]))
                yb = pin.to(dev, non_blocking=True)
                opt.zero_grad(set_to_none=True)
                mu, sig, ls = model(yb)
                loss = loss_fn(yb, mu, sig, ls)
                loss.backward()
                opt.step()
                return float(loss)
            step = None
        n = max(10, min(args.steps, 40))
        for i in range(3):
            run(i)
        barrier()
        t0 = time.perf_counter()
        if graphed and pipelined:
            # every step: pinned batch -> H2D -> graph replay -> loss D2H into pinned memory; the host reads a step's
            # loss a few steps later instead of stalling the launch of the next step on it
            tickets = [step.step_async(yd[((3 + i) * B) % 9984:((3 + i) * B) % 9984 + B]) for i in range(n)]
            last = step.loss_of(tickets[-1])
        else:
            for i in range(n):
                last = run(3 + i)
        barrier()
        dt = time.perf_counter() - t0
        return n, dt, last, step

    elbo_steps, elbo_s, last_loss_g, gstep = time_elbo(128 * world, 5)
    _, elbo_sync_s, _, gstep_sync = time_elbo(128 * world, 5, pipelined=False)
    del gstep_sync
    _, elbo_eager_s, last_loss, _ = time_elbo(128 * world, 5, graphed=False)
    elbo_nccl_s, nccl_loss = 0.0, None
    if use_peer:   # A/B: the same step with one NCCL all-reduce after the reduction kernel
        _, elbo_nccl_s, nccl_loss, gstep_nccl = time_elbo(128 * world, 5, peer=False)
        del gstep_nccl
        eng.peer_status()   # raises if any exchange timed out
    c3 = None
    if world == 1:  # config 3 (shapes of the shipped data file: S = 100)
        n3, c3_s, c3_loss, c3step = time_elbo(100, 3)
        c3 = {"steps_per_s": n3 / c3_s, "B": B, "S": 100, "samples_per_step": 6400, "last_loss": c3_loss,
              "fem_solves_per_s": 6400 * n3 / c3_s}
        del c3step

    # ---------------- config 4: Cook 80x40, batch 1024 per GPU (blocked panel kernel, factor streamed to HBM)
    md4 = pkg.PreProcessing.modeldata_initialization_topopt(pkg.cook_membrane_feap(80, 40))
    eng4 = pkg.CookFemEngine(md4, device=local, node_id=3321, ele_id=12)
    n4 = 1024
    x4 = torch.tensor(np.random.default_rng(4 + 100 * rank).standard_normal((n4, 2)), device=dev)
    g4 = np.random.default_rng(5 + 100 * rank).standard_normal((n4, 4))
    gy4, gh4 = torch.tensor(np.ascontiguousarray(g4[:, :2]), device=dev), torch.tensor(np.ascontiguousarray(g4[:, 2:]), device=dev)
    k4 = max(5, min(args.steps, 10))

    def time4(fn):
        for _ in range(3):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k4)]
        barrier()
        for ea, eb in evs:
            flush.zero_()
            ea.record()
            fn()
            eb.record()
        barrier()
        return float(sum(ea.elapsed_time(eb) for ea, eb in evs))
    c4_adj_ms = time4(lambda: eng4.forward_backward(x4, gy4, gh4))
    c4_bad = eng4.status(n4)[0]
    c4_fwd_ms = time4(lambda: eng4.forward(x4))
    c4_info = dict(eng4.info)

    # ---------------- max over ranks
    t = torch.tensor([dev_ms, b2b_ms, fwd_ms, e2e_s, elbo_s, elbo_eager_s, sus_ms, c4_adj_ms, c4_fwd_ms, elbo_sync_s,
                      elbo_nccl_s],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    (dev_ms, b2b_ms, fwd_ms, e2e_s, elbo_s, elbo_eager_s, sus_ms, c4_adj_ms, c4_fwd_ms, elbo_sync_s,
     elbo_nccl_s) = t.tolist()

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
        cpu = cpu_elbo = cpu4 = None
        if world == 1 and not args.no_cpu_baseline:
            fo, to, mesh, dof = oracle()
            ns = 1536
            cpu_fwd_adjoint(to, xh[:256], gyh[:256], ghh[:256])
            t0 = time.perf_counter()
            cpu_fwd_adjoint(to, xh[:ns], gyh[:ns], ghh[:ns])
            dt = time.perf_counter() - t0
            cpu = {"value": ns / dt, "unit": "solves/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"first {ns} of the 4096 seeded samples, forward+adjoint, oracle TorchOracle.vjp "
                             "(vectorised dense assembly + dense LU + autograd, torch CPU f64)",
                   "statement_by_statement_port": time_loop_port(),
                   "numpy_twin": time_numpy_twin(), "tf": tf_status()}
            cpu_elbo = cpu_elbo_step(to, fo)
            # config 4 on the CPU: the sparse port (CSR assembly + SuperLU + discrete adjoint), one core
            m4 = fo.read_mesh_text(fo.cook_mesh_text(80, 40))
            so = fo.SparseOracle(m4, fo.assign_dof(m4))
            xs4, gs4 = x4[:4].cpu().numpy(), g4[:4]
            so.vjp(xs4[:1], gs4[:1, :2], gs4[:1, 2:], 3321, 12)
            t0 = time.perf_counter()
            so.vjp(xs4, gs4[:, :2], gs4[:, 2:], 3321, 12)
            cpu4 = {"value": 4 / (time.perf_counter() - t0), "unit": "solves/s", "cores": 1, "kind": "port",
                    "sample": "4 seeded samples, forward+adjoint, oracle SparseOracle.vjp (CSR assembly + SuperLU "
                              "+ discrete adjoint), one core"}
        S5 = 128 * world
        elbo_rate = elbo_steps / elbo_s
        elbo_tflops = elbo_rate * B * S5 * FLOP_PER_SOLVE / 1e12
        c4_rate = world * n4 * k4 / (c4_adj_ms * 1e-3)
        c4_flop = 1600 * 3200 + 6560 * (85 * 85 + 3 * 85) + 2 * 4 * 6560 * 85       # 58.7 MFLOP
        c4_bytes = 4 * 6560 * 86 * 8 + 96                                            # 18.05 MB if L is streamed
        c4_tf = c4_rate / world * c4_flop / 1e12
        c4_gbs = c4_rate / world * c4_bytes / 1e9
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
                         "traffic_unit": "bytes of DRAM traffic per launch (ncu --set full, batch 4096, fused forward+adjoint): "
                                         "236 KB per sample = the scaled factor panels streamed to a per-warp slab and read "
                                         "back once by the reverse pass (SURVEY 8d counts 4 n (b+1) 8 = 366 KB per sample for "
                                         "a streamed factor); compulsory I/O is 96 B per sample",
                         "peak_source": "DFMA loop measured on this GPU by vbfem_measure_peaks "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_solve": FLOP_PER_SOLVE,
                         "hbm": {"achieved_gbs": BYTES_PER_SOLVE * BATCH / launch_s / 1e9, "peak_gbs": hbm_peak,
                                 "measured_dram_gbs": NCU_DRAM_BYTES_PER_LAUNCH / launch_s / 1e9,
                                 "note": "achieved_gbs = compulsory I/O only; measured_dram_gbs = the ncu traffic of one "
                                         "launch over this run's launch time (factor slab): well below the HBM peak, the "
                                         "kernel is bound by the FP64 / DMMA pipe and its dependency chains"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "elbo": {"metric": "ELBO training steps/s (Cook 20x10, 8192 Monte-Carlo samples per GPU)",
                     "value": elbo_rate, "unit": "steps/s", "n_gpus": world, "B": B, "S": S5,
                     "samples_per_step": B * S5, "samples_per_gpu": B * S5 // world, "scaling": "weak",
                     "fem_solves_per_s": B * S5 * elbo_rate, "last_loss": last_loss_g,
                     "cuda_graph": bool(gstep.graphed), "eager_steps_per_s": elbo_steps / elbo_eager_s,
                     "host_synchronised_every_step_steps_per_s": elbo_steps / elbo_sync_s,
                     "eager_last_loss": last_loss, "timed_steps": elbo_steps,
                     "collective": ("none (one GPU)" if world == 1 else
                                    "3 + 4B doubles per step exchanged INSIDE the timed region by the reduction kernel "
                                    "itself: P2P stores into every peer's mailbox over NVLink, sequence flags, sum in "
                                    "rank order (csrc/vbfem_peer.cuh; a node of the captured graph)" if use_peer else
                                    "one NCCL all-reduce of 3 + 4B doubles per step INSIDE the timed region (a node of "
                                    "the captured graph); peer mailboxes unavailable: " + str(peer_note)),
                     # A/B: range-restricted partials, one NCCL all-reduce node, KL terms and loss value in torch
                     "nccl_all_reduce_steps_per_s": (elbo_steps / elbo_nccl_s) if use_peer and elbo_nccl_s else None,
                     "nccl_last_loss": nccl_loss,
                     "roofline": {"bound": "fp64", "achieved": elbo_tflops, "peak": fp64.value * world,
                                  "unit": "TFLOP/s", "frac": elbo_tflops / (fp64.value * world) if fp64.value else None,
                                  "note": "0.716 MFLOP per reparameterised sample (FEM forward + adjoint); the two "
                                          "1884-parameter MLPs and Adam are < 0.1 % of the flops"},
                     "cpu_baseline": cpu_elbo,
                     "what": "every step: pinned batch H2D -> NN fwd -> [library: reparam -> FEM fwd -> data-term cotangent -> FEM adjoint -> reductions (+ the ranks' exchange) -> KL terms, loss] -> NN bwd "
                             "-> Adam (one CUDA graph replay) -> loss D2H into pinned memory; the host launches up to four "
                             "steps ahead and reads each loss afterwards (GraphedStep1.step_async); "
                             "host_synchronised_every_step_steps_per_s = the same with float(loss) after every step",
                     "config3": c3},
            "config4": {"workload": "Cook 80x40 (n=6560, half bandwidth 85), batch 1024 per GPU, fused forward+adjoint "
                                    "(BASELINE configs[3]); x seed 4, cotangents seed 5",
                        "value": c4_rate, "unit": "solves/s", "ms_per_step": c4_adj_ms / k4, "steps": k4,
                        "forward_only_solves_per_s": world * n4 * k4 / (c4_fwd_ms * 1e-3), "flagged_samples": c4_bad,
                        "l2": "flushed between timed steps", "kernel": c4_info,
                        "roofline": {"bound": "hbm if the factor is streamed (SURVEY 8d), fp64 otherwise",
                                     "fp64": {"achieved": c4_tf, "peak": fp64.value, "unit": "TFLOP/s",
                                              "frac": c4_tf / fp64.value if fp64.value else None,
                                              "flop_per_solve": c4_flop},
                                     "hbm": {"achieved": c4_gbs, "peak": hbm_peak, "unit": "GB/s",
                                             "frac": c4_gbs / hbm_peak, "bytes_per_solve": c4_bytes},
                                     "frac": min(c4_tf / fp64.value if fp64.value else 1.0, c4_gbs / hbm_peak),
                                     "traffic": NCU_C4_DRAM_BYTES_PER_SOLVE,
                                     "traffic_unit": "bytes of DRAM traffic per SOLVE (ncu --set full, 296 samples, "
                                                     "profiles/r02_ncu_80x40_adj_details.txt)"},
                        "cpu_baseline": cpu4},
            "latency": lat,
            "extra": {"back_to_back_solves_per_s": world * BATCH * args.steps / (b2b_ms * 1e-3),
                      "sustained": {"solves_per_s": world * BATCH * n_sus / (sus_ms * 1e-3), "launches": n_sus,
                                    "seconds": sus_ms * 1e-3, "what": "back-to-back fused launches, no L2 flush"},
                      "forward_only_solves_per_s": world * BATCH * args.steps / (fwd_ms * 1e-3),
                      "step_ms_min": min(step_ms), "step_ms_max": max(step_ms)},
        }
        print(json.dumps(line), flush=True)
    eng4.close()
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
