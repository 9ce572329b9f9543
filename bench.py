#!/usr/bin/env python
"""Headline benchmark: Grad_f evaluations per second (forward + discrete adjoint) of the kinematic dynamo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload kdyn128|kdyn64|kdyn24|sh23ens|sh23]
                    [--no-graph] [--no-cpu]

One "step" = one Grad_f evaluation of BASELINE config 3: f(X) followed by Grad_f(X) (the reference's state coupling:
Grad_f replays the snapshots f wrote), Npts = 128^3 (192^3 dealiased grid), Rm = 10, dt = 1e-3, N_ITERS = 1000 time
steps each way, X = [B0, U] synthetic band-limited solenoidal fields (seeded).  The line printed by rank 0 follows the
driver's contract; see DESIGN.md section "Measurement" for the definition of every key.

 * value  - device-resident vectors (DevVec), timed with CUDA events on the launching stream, max over ranks; the time
            loops are replayed from CUDA graphs captured during the warm-up (--no-graph: eager launches);
 * e2e    - the same pair through the reference-facing callables with HOST (pinned numpy) vectors in and numpy
            gradients out, H2D/D2H copies inside the timed region;
 * roofline - the dominant kernel (fused adjoint x-pass), CUDA-event timed per launch inside the timed region,
            algorithmic bytes per SURVEY.md section 8(d);
 * cpu_baseline - the numpy/scipy oracle (a port, not Dedalus) on the box's host cores, bounded sample.
Other workloads (parity-test configurations of BASELINE.json, not the headline): sh23ens = config 5 (4096 SH23 problems in one
batched launch each way), sh23 = config 1 (one SH23 problem), kdyn64 / kdyn24 = smaller dynamo grids.
With --impl reference the same oracle is the timed arm (the reference's own Dedalus path cannot be installed: no
dedalus/mpi4py/FFTW in the image and no network).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Npts, Rm, dt, N_ITERS)
    "kdyn128": (128, 10.0, 1e-3, 1000),
    "kdyn64": (64, 10.0, 1e-3, 1000),
    "kdyn24": (24, 1.0, 1e-3, 1000),
    # BASELINE config 5: 4096 independent SH23 problems (Npts=256, dt=0.1, T=50), M_0 swept over [0.05, 0.1], sharded over the GPUs
    "sh23ens": (256, None, 0.1, 500),
    # BASELINE config 1: ONE SH23 problem (the reference's own CPU-runnable case): latency of one f + Grad_f pair
    "sh23": (256, None, 0.1, 500),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
# (profiles/r1g_kdyn128_adj_step_ncu.txt: 509.6 MB read + 185.4 MB written - it also reads the forward state from its
# snapshot slot and read-modify-writes the running sum of the gradient integrand; algorithmic model 566.2 MB)
NCU_TRAFFIC = {("kdyn128", 1): 695.0e6}
METRIC = "Grad_f evals/s (fwd+adjoint)"
UNIT = "Grad_f evals/s"


def alg_bytes(N):
    """per scalar field C, P1, P2 bytes; forward step 9C+12P1+15P2, adjoint 18C+24P1+27P2"""
    M = 3 * N // 2
    Cb = (N // 2) * (N - 1) ** 2 * 16
    P1 = (N // 2) * (N - 1) * M * 16
    P2 = (N // 2) * M * M * 16
    return Cb, P1, P2, 9 * Cb + 12 * P1 + 15 * P2, 18 * Cb + 24 * P1 + 27 * P2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm) on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------------
def cpu_pair_sample(N, Rm, dt, n_iters_full, sample_steps):
    """time `sample_steps` forward + adjoint steps of the oracle at Npts = N; returns (evals/s extrapolated, info)"""
    import scipy.fft  # noqa: F401
    from oracle import fourier as ofo
    from oracle import kdyn as okd
    cores = os.cpu_count() or 1
    ofo.WORKERS = cores
    dom = okd.domain_kdyn(N)
    rng = np.random.RandomState(0)
    M = dom.M
    # cheap band-limited inputs: random coefficients with a spectral decay (no projection needed for timing)
    K = okd._K(dom)

    def field(seed):
        r = np.random.RandomState(seed)
        c = [(r.standard_normal(dom.coeff_shape) + 1j * r.standard_normal(dom.coeff_shape)) * np.exp(-0.3 * np.sqrt(K[3])) for _ in range(3)]
        c = okd._project(K, c)
        for ci in c:
            ci[0, :, :] = 0.0   # keep the kx = 0 plane trivially Hermitian
        return okd.Field_to_Vec(dom, *[dom.to_grid_3d(ci) for ci in c])
    B0, U = field(1), field(2)
    D = okd.GEN_BUFFER(N, dom, sample_steps)
    t0 = time.perf_counter()
    okd.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, sample_steps, sample_steps, D)
    t1 = time.perf_counter()
    okd.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, sample_steps, sample_steps, D)
    t2 = time.perf_counter()
    # per-step cost: subtract nothing (set-up transforms of U and the terminal J are included, which favours the GPU
    # arm by < 1 step); extrapolate linearly to n_iters_full steps each way
    per_step_pair = (t2 - t0) / sample_steps
    evals = 1.0 / (per_step_pair * n_iters_full)
    info = {"value": evals, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "numpy/scipy.fft oracle (port of the reference algorithm, not Dedalus), Npts=%d^3, %d of %d time steps "
                      "forward + adjoint with scipy.fft workers=%d: fwd %.2f s, adj %.2f s, extrapolated linearly"
                      % (N, sample_steps, n_iters_full, cores, t1 - t0, t2 - t1)}
    return evals, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, Rm, dt, nit = WORKLOADS[args.workload]
    if args.workload in ("sh23ens", "sh23"):
        from oracle import sh23 as osh
        od, X0 = osh.Generate_IC(0.0725)
        D = osh.GEN_BUFFER(od, nit)
        t0 = time.perf_counter()
        for _ in range(2):
            osh.FWD_Solve_IVP_Lin([X0], od, dt, nit, nit, D); osh.ADJ_Solve_IVP_Lin([X0], od, dt, nit, nit, D)
        value = 2.0 / (time.perf_counter() - t0)
        info = {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": "numpy oracle, 2 of 4096 instances, serial"}
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": (4096e3 if args.workload == "sh23ens" else 1e3) / value, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": "SH23 ensemble (config 5), oracle sample"},
                          "cpu_baseline": info, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    sample_steps = {128: 2, 64: 8, 24: 100}.get(N, 2)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue   # one warm-up sample is enough for a CPU arm (keeps the run within minutes)
        v, info = cpu_pair_sample(N, Rm, dt, nit, sample_steps)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, gpus):
    N, Rm, dt, nit = WORKLOADS[name]
    M = 3 * N // 2
    return {"workload": "kinematic dynamo Npts=%d^3 (grid %d^3), Rm=%g, dt=%g, N_ITERS=%d, cost Final, discrete adjoint; "
                        "one step = f(X) + Grad_f(X), X=[B0,U]" % (N, M, Rm, dt, nit),
            "Npts": N, "N_ITERS": nit, "dof": 3 * N ** 3, "decomposition": "z/kx slabs over %d GPU(s)" % gpus,
            "cache": "working set (x-spectral snapshot store %.1f GB over the GPUs + pencil fields) far larger than the 126 MB L2; no explicit flush"
                     % ((nit + 1) * 3 * (N // 2) * M * M * 16 / 1e9)}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from spheremanopt_b200 import _cabi, kdyn
    from spheremanopt_b200.devvec import DevVec
    lib = _cabi.load()
    N, Rm, dt, nit = WORKLOADS[args.workload]
    dom = kdyn.Domain(N, device="cuda:%d" % local)
    if not args.no_graph:
        lib.smo_kdyn_use_graph(dom.h, 1)     # time loops replayed from CUDA graphs (captured during warm-up)
    M = dom.M
    dev = dom.device

    # synthetic inputs: seeded noise -> band-limited through the library's own transforms -> unit norm
    def synth(seed):
        g = torch.Generator(device="cpu").manual_seed(seed)
        full = torch.randn(3, M, M, M, dtype=torch.float64, generator=g)
        slab = full[:, :, :, dom.z0:dom.z0 + dom.nz].contiguous().to(dev).reshape(-1)
        c = kdyn.to_coef(dom, slab)
        kx, ky, kz = kdyn._wavenumbers(dom)
        k2 = kx * kx + ky * ky + kz * kz
        c = c * torch.exp(-0.15 * torch.sqrt(k2))
        kdotc = (kx * c[0] + ky * c[1] + kz * c[2]) / torch.where(k2 == 0, torch.ones_like(k2), k2)
        c = torch.stack([c[0] - kx * kdotc, c[1] - ky * kdotc, c[2] - kz * kdotc]) * (k2 != 0)
        v = kdyn.to_grid(dom, c)
        return v / np.sqrt(kdyn.Inner_Prod_3(DevVec(v), DevVec(v), dom))
    B0, U = synth(1), synth(2)
    store = kdyn.GEN_BUFFER(N, dom, nit)
    Xd = [DevVec(B0), DevVec(U)]
    fargs = (dom, Rm, dt, nit, nit, store, "Final", "Discrete")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pair(X):
        f = kdyn.FWD_Solve_IVP_Lin(X, *fargs)
        g = kdyn.ADJ_Solve_IVP_Lin(X, *fargs)
        return f, g

    PK_XADJ = 6
    lib.smo_kdyn_profile_set(dom.h, PK_XADJ)     # before the warm-up: the event records become part of the captured graphs
    for _ in range(args.warmup):
        f, g = pair(Xd)
    # ---- timed region: K pairs, device-resident inputs ------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    lib.smo_kdyn_profile_set(dom.h, PK_XADJ)
    n0 = lib.smo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        f, g = pair(Xd)
    e1.record()
    barrier()
    launches = lib.smo_launch_count() - n0
    ms_total = e0.elapsed_time(e1)
    kms, kn = C.c_double(), C.c_longlong()
    lib.smo_kdyn_profile_read(dom.h, C.byref(kms), C.byref(kn))
    lib.smo_kdyn_profile_set(dom.h, 0)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = 1e3 / ms_step
    # size-independent checks of the result (identical for every GPU count: compare the lines of a scaling run)
    gnorm = [kdyn.Inner_Prod_3(g[0], g[0], dom), kdyn.Inner_Prod_3(g[1], g[1], dom)]

    # ---- e2e: host vectors in, host gradients out (reference-facing Mode H) ----------------------------------
    Bh = torch.empty(3 * M ** 3, dtype=torch.float64, pin_memory=True)
    Uh = torch.empty(3 * M ** 3, dtype=torch.float64, pin_memory=True)
    Bh.copy_(torch.from_numpy(dom.host_from_slab(B0))); Uh.copy_(torch.from_numpy(dom.host_from_slab(U)))
    Xh = [Bh.numpy(), Uh.numpy()]
    e2e_steps = max(1, min(args.steps, 2))
    w1 = pair(Xh); w2 = pair(Xh)   # warm the host path: two generations of page-locked result buffers enter torch's host
    del w1, w2                     # allocator cache (the timed loop keeps one generation alive while it fills the next)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fh, gh = pair(Xh)
    barrier()
    t1 = time.perf_counter()
    te = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = 1.0 / float(te.item())
    h2d = 2 * 3 * M * M * dom.nz * 8 * world
    d2h = 2 * 3 * M ** 3 * 8 * world   # every rank receives the full gradients, like the reference's allgather
    assert isinstance(gh[0], np.ndarray) and abs(fh - f) <= 1e-9 * abs(f)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    Cb, P1, P2, af, aa = alg_bytes(N)
    peak, peak_src = peaks()
    k_alg = 15 * P2 / world                    # fused adjoint x-pass: 9 P2 read + 6 P2 written
    k_ms = kms.value / max(kn.value, 1)
    ach = k_alg / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    pair_gbs = (af + aa) * nit / world / (ms_step * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.workload, world),
            "dof_steps_per_s": 3 * N ** 3 * 2 * nit * value,
            "J": -f, "grad_norms": gnorm,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "XPass<X_ADJ> (fused c2r + (curl G)xU, (curl G)xB_f + r2c, adjoint step)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "traffic": NCU_TRAFFIC.get((args.workload, world)),
                         "launch_ms": k_ms, "launches_timed": int(kn.value), "share_of_step": kms.value / (ms_step * args.steps),
                         "algorithmic_bytes_per_launch": k_alg, "peak_source": peak_src},
            "roofline_pair": {"bound": "hbm", "achieved": pair_gbs, "peak": peak, "unit": "GB/s", "frac": pair_gbs / peak,
                              "algorithmic_bytes_per_pair_per_gpu": (af + aa) * nit / world,
                              "note": "whole Grad_f pair, all kernels and launch gaps; per-GPU algorithmic bytes / wall"}}
    if world == 1 and not args.no_cpu:
        try:
            _, info = cpu_pair_sample(N, Rm, dt, nit, {128: 2, 64: 8, 24: 100}.get(N, 2))
            line["cpu_baseline"] = info
        except Exception as e:   # the baseline is a report, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_gpu_sh23ens(args):
    """config 5: one step = f + Grad_f of the whole ensemble (4096 instances, one kernel launch each way per GPU)"""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from spheremanopt_b200 import _cabi, sh23
    lib = _cabi.load()
    N, _, dt, nit = WORKLOADS[args.workload]
    total = 4096 if args.workload == "sh23ens" else world      # config 1: one instance (per GPU: replicas only)
    nb = total // world
    dom, X0 = sh23.Generate_IC(0.0725, N, device="cuda:%d" % local)
    M0 = (np.linspace(0.05, 0.1, total) if args.workload == "sh23ens" else np.full(total, 0.0725))[rank * nb:(rank + 1) * nb]
    X = torch.from_numpy(np.sqrt(M0 / 0.0725)[:, None] * X0[None, :]).to(dom.device).reshape(-1).contiguous()
    store = sh23.GEN_BUFFER(dom, nit, N, batch=nb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pair(x):
        J = sh23.forward_batch(x, dom, dt, nit, store)
        G = sh23.adjoint_batch(dom, dt, nit, store)
        return J, G
    for _ in range(args.warmup):
        pair(X)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = lib.smo_launch_count()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    tf = ta = 0.0
    for _ in range(args.steps):
        e[1].record(); J = sh23.forward_batch(X, dom, dt, nit, store)
        e[2].record(); G = sh23.adjoint_batch(dom, dt, nit, store)
        e[3].record()
    e3 = torch.cuda.Event(enable_timing=True); e3.record()
    barrier()
    launches = lib.smo_launch_count() - n0
    ms_total = e[0].elapsed_time(e3)
    t_adj = e[2].elapsed_time(e[3])     # last adjoint launch (the dominant kernel)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dom.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = total / (ms_step * 1e-3)
    # e2e: pinned host vectors in, host gradients + J out
    Xh = torch.empty(nb * dom.M, dtype=torch.float64, pin_memory=True); Xh.copy_(X.cpu())
    Gh = torch.empty(nb * dom.M, dtype=torch.float64, pin_memory=True)
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        Xd = Xh.to(dom.device, non_blocking=True)
        J, G = pair(Xd)
        Gh.copy_(G, non_blocking=True); Jh = J.cpu()
    barrier(); t1 = time.perf_counter()
    te = torch.tensor([(t1 - t0) / args.steps], dtype=torch.float64, device=dom.device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    alg_inst = 2 * (nit + 1) * (N // 2) * 16 + 2 * 2 * N * 8          # snapshots written + read, X in, gradient out
    alg_adj = nb * ((nit + 1) * (N // 2) * 16 + 2 * N * 8)
    ach = alg_adj / (t_adj * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("SH23 ensemble: %d independent problems" % total if total > world else "SH23 single problem (config 1; replicas only on > 1 GPU)")
                                   + " Npts=%d (grid %d), dt=%g, N_ITERS=%d, M_0 in [0.05,0.1], discrete adjoint; "
                                   "one step = f + Grad_f of every instance; sharded %d per GPU, no collective" % (N, 2 * N, dt, nit, nb),
                       "instances": total, "Npts": N, "N_ITERS": nit, "dof": N,
                       "cache": "snapshot store %.1f GB per GPU streams through HBM; the state of an instance lives in shared memory" % (nb * (nit + 1) * (N // 2) * 16 / 1e9)},
            "dof_steps_per_s": N * 2 * nit * value,
            "e2e": {"value": total / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": nb * dom.M * 8 * world, "d2h_bytes_per_step": (nb * dom.M * 8 + nb * 8) * world},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "Sh23Adj (whole adjoint time loop of %d instances, one launch)" % nb, "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "launch_ms": t_adj, "algorithmic_bytes_per_launch": alg_adj,
                         "peak_source": peak_src,
                         "note": "serial fp64 chain per instance: the fp64 pipe and shared memory bind, not HBM (SURVEY 8(d)); %d B algorithmic per instance pair" % alg_inst}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kdyn128", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the time loops eagerly (no CUDA graphs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("sh23ens", "sh23"):
        run_gpu_sh23ens(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
