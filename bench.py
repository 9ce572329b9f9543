#!/usr/bin/env python
"""Headline benchmark: Grad_f evaluations per second (forward + discrete adjoint) of the kinematic dynamo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload kdyn128|kdyn256|kdyn64|kdyn24|sh23ens|sh23|vec] [--no-graph] [--no-cpu]

One "step" = one Grad_f evaluation of BASELINE config 3: f(X) followed by Grad_f(X) (the reference's state coupling:
Grad_f replays the snapshots f wrote), Npts = 128^3 (192^3 dealiased grid), Rm = 10, dt = 1e-3, N_ITERS = 1000 time
steps each way, X = [B0, U] synthetic band-limited solenoidal fields (seeded).  The line printed by rank 0 follows the
driver's contract; see DESIGN.md section "Measurement" for the definition of every key.

 * value  - device-resident vectors (DevVec), timed with CUDA events on the launching stream, max over ranks; the time
            loops are replayed from CUDA graphs captured during the warm-up (--no-graph: eager launches);
 * e2e    - the same pair through the reference-facing callables with HOST (pinned numpy) vectors in and numpy
            gradients out, H2D/D2H copies inside the timed region, over the same number of steps; the copy times
            (h2d_ms / d2h_ms per step, CUDA events around copies of the same buffers) are printed next to it;
 * roofline - the dominant kernel (fused adjoint x-pass), CUDA-event timed per launch (event-record nodes inside the replayed
            graphs) over a SECOND pass of the same K steps right after the headline region - the event nodes cost a few us
            per launch, which round 1 had charged to `value`; algorithmic bytes per SURVEY.md section 8(d);
 * cpu_baseline - the numpy/scipy oracle (a port, not Dedalus) on the box's host cores, bounded sample;
 * mp_parity_relerr (N > 1) - worst relative error of J / Grad_f of a 32^3, 6-step run on the SAME process group against
            the oracle, taken before the timed region: multi-GPU parity evidence inside the scaling record.
Other workloads (parity-test configurations of BASELINE.json, not the headline): kdyn256 = config 4 (256^3, Rm = 20,
checkpointed adjoint when the snapshots do not fit), kdyn24 = config 2, sh23ens = config 5 (4096 SH23 problems in one
batched launch each way), sh23 = config 1 (one SH23 problem; with the reference optimiser staged under baseline/_ref also
the full optimisation), vec = the inner-product / sphere-geometry kernels (rows A5/B5, C1-C3) against the HBM peak.
With --impl reference the oracle is the timed arm (the reference's own Dedalus path cannot be installed: no
dedalus/mpi4py/FFTW in the image and no network); its samples are extrapolated and say so.
"""
import os
import sys

if "reference" in sys.argv:
    # the CPU arm uses every host core whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1): set before numpy loads
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse      # noqa: E402
import ctypes as C   # noqa: E402
import json          # noqa: E402
import subprocess    # noqa: E402
import threading     # noqa: E402
import time          # noqa: E402

import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Npts, Rm, dt, N_ITERS)
    "kdyn128": (128, 10.0, 1e-3, 1000),      # BASELINE config 3 (the headline)
    "kdyn256": (256, 20.0, 1e-3, 1000),      # BASELINE config 4: checkpointed adjoint where the 907 GB of snapshots do not fit
    "kdyn64": (64, 10.0, 1e-3, 1000),
    "kdyn24": (24, 1.0, 1e-3, 1000),         # BASELINE config 2 (latency bound: working set < L2)
    # BASELINE config 5: 4096 independent SH23 problems (Npts=256, dt=0.1, T=50), M_0 swept over [0.05, 0.1], sharded over the GPUs
    "sh23ens": (256, None, 0.1, 500),
    # BASELINE config 1: ONE SH23 problem (the reference's own CPU-runnable case): latency of one f + Grad_f pair
    "sh23": (256, None, 0.1, 500),
    # BASELINE config 5 as stated: 4096 independent OPTIMISATIONS (unmodified reference optimiser, needs baseline/_ref), 512 per GPU,
    # every f / Grad_f / Inner_Product call of the ensemble served by batched launches (spheremanopt_b200/ensemble.py)
    "sh23opt": (256, None, 0.1, 500),
    # rows A5/B5, C1-C3: vector kernels on dynamo-sized vectors (3 * 192^3 doubles = 170 MB each, > L2)
    "vec": (128, None, None, None),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
# named in NCU_SOURCE (it also reads the forward state from its snapshot slot and read-modify-writes the running sum of the
# gradient integrand on the real grid - 340 MB that replace three r2c transforms; algorithmic model 566.2 MB)
NCU_TRAFFIC = {("kdyn128", 1): 803.4e6}
NCU_SOURCE = {("kdyn128", 1): "ncu --set full, build r2l (profiles/r2l_kdyn128_xpass_ncu.txt): 566.3 MB read + 237.1 MB written"}
METRIC = "Grad_f evals/s (fwd+adjoint)"
UNIT = "Grad_f evals/s"
# fp64 work of one SH23 instance pair (~138 GFLOP per 4096-instance pair at N_ITERS = 500): per time step one
# inverse + one forward (f) and two inverse + one forward (Grad_f) complex FFTs of length 256 at 5 N log2 N flop, the even/odd
# pre/post-processing (~10 flop per mode) and the pointwise terms
SH23_FLOP_PER_STEP_PAIR = 5 * (5 * 256 * 8) + 5 * 10 * 256 + 14 * 512


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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
_CPU_INPUTS = {}


def cpu_pair_sample(N, Rm, dt, n_iters_full, sample_steps):
    """time `sample_steps` forward + adjoint steps of the oracle at Npts = N; returns (evals/s extrapolated, info)"""
    import scipy.fft  # noqa: F401
    from oracle import fourier as ofo
    from oracle import kdyn as okd
    cores = host_cores()
    ofo.WORKERS = cores
    if N not in _CPU_INPUTS:
        dom = okd.domain_kdyn(N)
        K = okd._K(dom)

        def field(seed):     # cheap band-limited inputs: random coefficients with a spectral decay
            r = np.random.RandomState(seed)
            c = [(r.standard_normal(dom.coeff_shape) + 1j * r.standard_normal(dom.coeff_shape)) * np.exp(-0.3 * np.sqrt(K[3])) for _ in range(3)]
            c = okd._project(K, c)
            for ci in c:
                ci[0, :, :] = 0.0   # keep the kx = 0 plane trivially Hermitian
            return okd.Field_to_Vec(dom, *[dom.to_grid_3d(ci) for ci in c])
        _CPU_INPUTS[N] = (dom, field(1), field(2))
    dom, B0, U = _CPU_INPUTS[N]
    D = okd.GEN_BUFFER(N, dom, sample_steps)
    t0 = time.perf_counter()
    okd.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, sample_steps, sample_steps, D)
    t1 = time.perf_counter()
    okd.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, sample_steps, sample_steps, D)
    t2 = time.perf_counter()
    # linear extrapolation to n_iters_full steps each way; the set-up (projection of U, terminal J, final transforms) is
    # inside the sample, which overstates the per-step cost by (set-up / sample_steps) and so favours the GPU arm slightly
    per_step_pair = (t2 - t0) / sample_steps
    evals = 1.0 / (per_step_pair * n_iters_full)
    info = {"value": evals, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True, "sample_steps": sample_steps,
            "sample_seconds": t2 - t0,
            "sample": "numpy/scipy.fft oracle (port of the reference algorithm, not Dedalus), Npts=%d^3, %d of %d time steps "
                      "forward + adjoint with scipy.fft workers=%d: fwd %.2f s, adj %.2f s, extrapolated linearly"
                      % (N, sample_steps, n_iters_full, cores, t1 - t0, t2 - t1)}
    return evals, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, Rm, dt, nit = WORKLOADS[args.workload]
    if args.workload == "vec":
        n = 3 * 192 ** 3
        r = np.random.RandomState(0)
        x, y = r.standard_normal(n), r.standard_normal(n)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            float(np.dot(x, y))
        value = args.steps * 2 * n * 8 / (time.perf_counter() - t0) / 1e9
        print(json.dumps({"impl": "reference", "metric": "Inner_Product GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 2 * n * 8 / value / 1e6, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": "vector kernels: numpy dot of two 3*192^3 vectors"},
                          "cpu_baseline": {"value": value, "unit": "GB/s", "cores": host_cores(), "kind": "port", "sample": "np.dot, %d repetitions" % args.steps},
                          "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if args.workload in ("sh23ens", "sh23"):
        from oracle import sh23 as osh
        od, X0 = osh.Generate_IC(0.0725)
        D = osh.GEN_BUFFER(od, nit)
        nsamp = 4
        t0 = time.perf_counter()
        for _ in range(nsamp):
            osh.FWD_Solve_IVP_Lin([X0], od, dt, nit, nit, D); osh.ADJ_Solve_IVP_Lin([X0], od, dt, nit, nit, D)
        value = nsamp / (time.perf_counter() - t0)       # instance pairs per second, one core
        info = {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "extrapolated": args.workload == "sh23ens",
                "sample": "numpy oracle, %d instance pairs run serially on one core" % nsamp}
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": (4096e3 if args.workload == "sh23ens" else 1e3) / value, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": "SH23 (config %s), oracle sample" % ("5" if args.workload == "sh23ens" else "1")},
                          "cpu_baseline": info, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    # one warm-up sample (>= 10 time steps where that stays bounded) fixes the cost of a time step; the timed samples then use as
    # many steps (2..10) as keep the whole run within ~3 minutes, and are extrapolated linearly to N_ITERS
    first = {256: 2, 128: 10, 64: 20, 24: 200}.get(N, 2)
    budget_s = 150.0
    _, winfo = cpu_pair_sample(N, Rm, dt, nit, first)
    per_step = winfo["sample_seconds"] / first
    sample_steps = int(max(2, min(first, budget_s / max(args.steps, 1) / per_step)))
    vals, info = [], None
    for _ in range(args.steps):
        v, info = cpu_pair_sample(N, Rm, dt, nit, sample_steps)
        vals.append(v)
    value = float(np.mean(vals))
    info["value"] = value
    info["warmup_sample"] = "%d steps in %.1f s" % (first, winfo["sample_seconds"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "extrapolated": True, "sample_steps": sample_steps,
            "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, gpus, store=None):
    N, Rm, dt, nit = WORKLOADS[name]
    M = 3 * N // 2
    cfg = {"workload": "kinematic dynamo Npts=%d^3 (grid %d^3), Rm=%g, dt=%g, N_ITERS=%d, cost Final, discrete adjoint; "
                       "one step = f(X) + Grad_f(X), X=[B0,U]" % (N, M, Rm, dt, nit),
           "Npts": N, "N_ITERS": nit, "dof": 3 * N ** 3, "decomposition": "z/kx slabs over %d GPU(s)" % gpus,
           "cache": "working set (x-spectral snapshots: %.1f GB for all %d states over the GPUs, + pencil fields) far larger than the 126 MB L2; no explicit flush"
                    % ((nit + 1) * 3 * (N // 2) * M * M * 16 / 1e9, nit + 1)}
    if store is not None:
        cfg.update(store)
    return cfg


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def _init_dist():
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
    return torch, dist, world, rank, local


def mp_parity_check(kdyn, local):
    """multi-GPU parity evidence for the scaling record: 32^3, 6 steps on the same process group against the oracle"""
    from oracle import kdyn as okd
    from tests.common import kdyn_field, relerr
    Npts, nit = 32, 6
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 1.0, 1e-3, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 1.0, 1e-3, nit, nit, D)
    dom = kdyn.Domain(Npts, device="cuda:%d" % local)
    st = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, st)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, st)
    err = max(abs(f - fo) / abs(fo), relerr(g[0], go[0]), relerr(g[1], go[1]))
    del dom, st
    return float(err)


def run_gpu(args):
    torch, dist, world, rank, local = _init_dist()
    from spheremanopt_b200 import _cabi, kdyn
    from spheremanopt_b200.devvec import DevVec
    lib = _cabi.load()
    N, Rm, dt, nit = WORKLOADS[args.workload]
    mp_err = None
    if world > 1 and (N // 2) % world == 0:
        try:
            mp_err = mp_parity_check(kdyn, local) if (16 % world == 0 and 48 % world == 0) else None
        except Exception as e:      # evidence only: never lose the bench line over it
            mp_err = "failed: %r" % (e,)
    dom = kdyn.Domain(N, device="cuda:%d" % local)
    if not args.no_graph:
        lib.smo_kdyn_use_graph(dom.h, 1)     # time loops replayed from CUDA graphs (captured during warm-up)
    for k, v in os.environ.items():          # development switches (tools/): SMO_OPT_<key>=<value>, CHUNKS=f,a
        if k.startswith("SMO_OPT_"):
            lib.smo_kdyn_set_option(dom.h, int(k[8:]), int(v))
    if os.environ.get("CHUNKS"):
        cf, ca = (int(v) for v in os.environ["CHUNKS"].split(","))
        lib.smo_kdyn_set_chunks(dom.h, cf, ca)
    M = dom.M
    dev = dom.device

    # synthetic inputs: seeded noise -> band-limited through the library's own transforms -> unit norm
    def synth(seed):
        g = torch.Generator(device="cpu").manual_seed(seed)
        slab = torch.empty(3, M, M, dom.nz, dtype=torch.float64)
        for c in range(3):        # (plane by plane: the full 3 x M^3 noise field of a 256^3 run would be 1.4 GB per rank)
            full = torch.randn(M, M, M, dtype=torch.float64, generator=g)
            slab[c] = full[:, :, dom.z0:dom.z0 + dom.nz]
        slab = slab.to(dev).reshape(-1)
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
    ckpt = isinstance(store, kdyn.CheckpointStore)
    store_info = {"store": ("two-level checkpoints (revolve style): coefficients of every %d-th state + one recomputed x-spectral segment" % store.every)
                  if ckpt else "all %d states in HBM as x-spectra" % (nit + 1),
                  "checkpoint_every": store.every if ckpt else 0,
                  "states_held": store.states_held if ckpt else nit + 1,
                  "rho": store.rho if ckpt else 0.0,
                  "store_gb_per_gpu": (store.buf.numel() + (store.seg.numel() if ckpt else 0)) * 16 / 1e9}
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
    for _ in range(args.warmup):
        f, g = pair(Xd)
    # ---- timed region: K pairs, device-resident inputs ------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = lib.smo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        f, g = pair(Xd)
    e1.record()
    barrier()
    launches = lib.smo_launch_count() - n0
    ms_total = e0.elapsed_time(e1)
    # ---- the same K pairs once more with a CUDA-event pair around every launch of the dominant kernel (event-record nodes in
    # the replayed graphs).  Kept out of the headline region: the event nodes cost a few us per launch (r2a: 2 % of the pair at
    # 128^3, 30 % at 24^3).  roofline.share_of_step is this kernel's time over THIS pass's own wall time.
    lib.smo_kdyn_profile_set(dom.h, PK_XADJ)
    for _ in range(2):
        pair(Xd)                                 # (a new profile kind is a new graph key: eager run, then capture)
    lib.smo_kdyn_profile_set(dom.h, PK_XADJ)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        pair(Xd)
    p1.record()
    barrier()
    prof_ms_total = p0.elapsed_time(p1)
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

    # ---- e2e: host vectors in, host gradients out (reference-facing Mode H), same number of steps ------------
    Bh = torch.empty(3 * M ** 3, dtype=torch.float64, pin_memory=True)
    Uh = torch.empty(3 * M ** 3, dtype=torch.float64, pin_memory=True)
    Bh.copy_(torch.from_numpy(dom.host_from_slab(B0))); Uh.copy_(torch.from_numpy(dom.host_from_slab(U)))
    Xh = [Bh.numpy(), Uh.numpy()]
    e2e_steps = max(1, args.steps)
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
    # the copies of one step on their own (CUDA events): what the e2e figure adds to `value`
    c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    c0.record()
    s1, s2 = dom.slab_from_host(Xh[0]), dom.slab_from_host(Xh[1])
    c1.record()
    o1, o2 = dom.host_from_slab(g[0].t), dom.host_from_slab(g[1].t)
    c2.record()
    torch.cuda.synchronize()
    h2d_ms, d2h_ms = c0.elapsed_time(c1), c1.elapsed_time(c2)
    del s1, s2, o1, o2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    Cb, P1, P2, af, aa = alg_bytes(N)
    peak, peak_src = peaks()
    k_alg = 15 * P2 / world                    # fused adjoint x-pass: 9 P2 read + 6 P2 written
    k_ms = kms.value / max(kn.value, 1)
    ach = k_alg / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    rho = store_info["rho"]
    pair_bytes = (af * (1.0 + rho) + aa) * nit / world      # a checkpointed sweep recomputes rho forward solves
    pair_gbs = pair_bytes / (ms_step * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.workload, world, store_info),
            "dof_steps_per_s": 3 * N ** 3 * 2 * nit * value,
            "J": -f, "grad_norms": gnorm,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": 1e3 / e2e_val, "h2d_ms": h2d_ms, "d2h_ms": d2h_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "XFused<X_ADJ> (fused c2r + (curl G)xU, (curl G)xB_f + r2c + gradient-integrand accumulation, adjoint step)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "traffic": NCU_TRAFFIC.get((args.workload, world)), "traffic_source": NCU_SOURCE.get((args.workload, world)),
                         "launch_ms": k_ms, "launches_timed": int(kn.value), "share_of_step": kms.value / prof_ms_total,
                         "timed_in": "second pass of the same %d steps with event-record nodes around this kernel (%.1f ms per step in that pass)" % (args.steps, prof_ms_total / args.steps),
                         "algorithmic_bytes_per_launch": k_alg, "peak_source": peak_src},
            "roofline_pair": {"bound": "hbm", "achieved": pair_gbs, "peak": peak, "unit": "GB/s", "frac": pair_gbs / peak,
                              "algorithmic_bytes_per_pair_per_gpu": pair_bytes,
                              "note": "whole Grad_f pair, all kernels and launch gaps; per-GPU algorithmic bytes (SURVEY 8(d) model, forward part x (1 + rho)) / wall"}}
    if mp_err is not None:
        line["mp_parity_relerr"] = mp_err
    if world == 1 and not args.no_cpu:
        try:
            _, info = cpu_pair_sample(N, Rm, dt, nit, {256: 1, 128: 4, 64: 8, 24: 100}.get(N, 2))
            line["cpu_baseline"] = info
        except Exception as e:   # the baseline is a report, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_dfma_peak(torch, lib):
    """measured fp64 FMA throughput of this GPU (TFLOP/s): dependent-chain microbenchmark of the library, best of 5"""
    out = torch.zeros(256 * 8 * 256, dtype=torch.float64, device="cuda")
    flops = C.c_double()
    best = 0.0
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.smo_microbench_dfma(out.data_ptr(), 20000, 8, C.byref(flops), torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_gpu_sh23ens(args):
    """config 5: one step = f + Grad_f of the whole ensemble (4096 instances, one kernel launch each way per GPU)"""
    torch, dist, world, rank, local = _init_dist()
    from spheremanopt_b200 import _cabi, sh23
    lib = _cabi.load()
    N, _, dt, nit = WORKLOADS[args.workload]
    from spheremanopt_b200.ensemble import shard_slice
    total = 4096 if args.workload == "sh23ens" else world      # config 1: one instance (per GPU: replicas only)
    lo, hi = shard_slice(total, world, rank)
    nb = hi - lo
    dom, X0 = sh23.Generate_IC(0.0725, N, device="cuda:%d" % local)
    M0 = (np.linspace(0.05, 0.1, total) if args.workload == "sh23ens" else np.full(total, 0.0725))[lo:hi]
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
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for s in range(args.steps):
        ev[s][0].record(); J = sh23.forward_batch(X, dom, dt, nit, store)
        ev[s][1].record(); G = sh23.adjoint_batch(dom, dt, nit, store)
        ev[s][2].record()
    barrier()
    launches = lib.smo_launch_count() - n0
    ms_total = ev[0][0].elapsed_time(ev[-1][2])
    t_fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t_adj = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))     # the dominant kernel
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
    dfma = measure_dfma_peak(torch, lib) if rank == 0 else None
    opt = None
    if rank == 0 and args.workload == "sh23":
        opt = sh23_full_optimisation(sh23)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    alg_inst = 2 * (nit + 1) * (N // 2) * 16 + 2 * 2 * N * 8          # snapshots written + read, X in, gradient out
    alg_adj = nb * ((nit + 1) * (N // 2) * 16 + 2 * N * 8)
    ach = alg_adj / (t_adj * 1e-3) / 1e9
    flop_adj = nb * nit * (3 * (5 * 256 * 8) + 3 * 10 * 256 + 10 * 512)         # adjoint: 2 inverse + 1 forward FFT per step
    flop_pair = nb * nit * SH23_FLOP_PER_STEP_PAIR
    tf_adj = flop_adj / (t_adj * 1e-3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("SH23 ensemble: %d independent problems" % total if total > world else "SH23 single problem (config 1; replicas only on > 1 GPU)")
                                   + " Npts=%d (grid %d), dt=%g, N_ITERS=%d, M_0 in [0.05,0.1], discrete adjoint; "
                                   "one step = f + Grad_f of every instance; sharded %d per GPU, no collective" % (N, 2 * N, dt, nit, nb),
                       "instances": total, "Npts": N, "N_ITERS": nit, "dof": N,
                       "cache": "snapshot store %.1f GB per GPU streams through HBM; the state of an instance lives in shared memory" % (nb * (nit + 1) * (N // 2) * 16 / 1e9)},
            "dof_steps_per_s": N * 2 * nit * value,
            "ms_forward": t_fwd, "ms_adjoint": t_adj,
            "e2e": {"value": total / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": nb * dom.M * 8 * world, "d2h_bytes_per_step": (nb * dom.M * 8 + nb * 8) * world},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "Sh23Adj (whole adjoint time loop of %d instances, one launch)" % nb, "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "launch_ms": t_adj, "algorithmic_bytes_per_launch": alg_adj,
                         "peak_source": peak_src,
                         "note": "serial fp64 chain per instance: the fp64 pipe and shared memory bind, not HBM (SURVEY 8(d)); %d B algorithmic per instance pair" % alg_inst},
            "roofline_fp64": {"bound": "fp64 FMA pipe", "kernel": "Sh23Adj", "achieved": tf_adj, "peak": dfma, "unit": "TFLOP/s", "frac": (tf_adj / dfma) if dfma else None,
                              "pair_achieved": flop_pair / (ms_step * 1e-3) / 1e12, "algorithmic_flop_per_launch": flop_adj,
                              "peak_source": "measured in this run: smo_microbench_dfma (dependent DFMA chains, 8 CTAs x 256 threads per SM)",
                              "note": "algorithmic flop = 5 N log2 N per length-256 complex FFT + pre/post-processing + pointwise terms (SURVEY 8(d))"}}
    if opt is not None:
        line["optimisation"] = opt
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sh23_full_optimisation(sh23):
    """config 1 end to end: the UNMODIFIED reference optimiser (staged under baseline/_ref, tools/stage_reference.py) driving the
    CUDA callables with SH:783's arguments.  Returns None when the optimiser files are not there."""
    import tempfile
    ref = None
    for d in (os.environ.get("SMO_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "Sphere_Grad_Descent.py")):
            ref = d
            break
    if ref is None:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests", "refstubs")); sys.path.insert(0, ref)
    import Sphere_Grad_Descent as SGD
    E_0, nit = 0.0725, 500
    dom, X0 = sh23.Generate_IC(E_0)
    store = sh23.GEN_BUFFER(dom, nit)
    calls = {"f": 0, "g": 0, "ip": 0}

    def f(X, *a):
        calls["f"] += 1
        return sh23.FWD_Solve_IVP_Lin(X, *a)

    def g(X, *a):
        calls["g"] += 1
        return sh23.ADJ_Solve_IVP_Lin(X, *a)

    def ip(x, y, *a):
        calls["ip"] += 1
        return sh23.Inner_Prod(x, y, *a)
    import contextlib
    import warnings
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with contextlib.redirect_stdout(sys.stderr), warnings.catch_warnings():     # (the optimiser prints; stdout carries ONE JSON line)
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            RES, FUN, Xopt = SGD.Optimise_On_Multi_Sphere([X0], [E_0], f, g, ip, [dom, 0.1, nit, nit, store, None, "Discrete"], (dom, None),
                                                          max_iters=200, alpha_k=np.pi, LS='LS_wolfe', CG=True, callback=None, verbose=False)
            sec = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    return {"what": "Optimise_On_Multi_Sphere (unmodified, %s) with SH:783's arguments on the CUDA callables, host vectors" % ref,
            "seconds": sec, "iterations": len(FUN), "f_calls": calls["f"], "grad_calls": calls["g"], "inner_product_calls": calls["ip"],
            "J_final": float(FUN[-1]), "residual_final": float(RES[0][-1])}


def run_gpu_sh23opt(args):
    """config 5 as stated: 4096 / world independent optimisations per GPU (M_0 swept over [0.05, 0.1]), the UNMODIFIED reference
    optimiser in one thread per instance, all calls of one kind served by ONE batched launch (ensemble.SH23Ensemble).
    One step = the whole ensemble optimised from its initial fields (max_iters = $SMO_ENS_ITERS, default 20)."""
    import contextlib
    import tempfile
    import warnings
    torch, dist, world, rank, local = _init_dist()
    from spheremanopt_b200 import _cabi, sh23
    from spheremanopt_b200.ensemble import SH23Ensemble
    lib = _cabi.load()
    ref = None
    for d in (os.environ.get("SMO_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "Sphere_Grad_Descent.py")):
            ref = d
            break
    if ref is None:
        if rank == 0:
            print(json.dumps({"metric": "SH23 optimisations/s", "unavailable": "reference optimiser not staged (tools/stage_reference.py)"}))
        return
    sys.path.insert(0, os.path.join(ROOT, "tests", "refstubs")); sys.path.insert(0, ref)
    import Sphere_Grad_Descent as SGD
    N, _, dt, nit = WORKLOADS[args.workload]
    total = int(os.environ.get("SMO_ENS_TOTAL", "4096"))
    iters = int(os.environ.get("SMO_ENS_ITERS", "20"))
    from spheremanopt_b200.ensemble import shard_slice
    lo, hi = shard_slice(total, world, rank)
    nb = hi - lo
    dom, X0 = sh23.Generate_IC(0.0725, N, device="cuda:%d" % local)
    M0 = np.linspace(0.05, 0.1, total)[lo:hi]
    X0s = [np.sqrt(m / 0.0725) * X0 for m in M0]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    times, last = [], None
    n0 = lib.smo_launch_count()
    try:
        for s_ in range(args.warmup + args.steps):
            ens = SH23Ensemble(nb, dom, dt, nit)

            def one(i, f, g, ip):
                try:
                    return SGD.Optimise_On_Multi_Sphere([X0s[i]], [float(M0[i])], f, g, ip, [dom, dt, nit, nit, None, None, "Discrete"], (dom, None),
                                                        max_iters=iters, alpha_k=np.pi, LS='LS_wolfe', CG=True, callback=None, verbose=False)
                except TypeError:
                    # the reference's own failure mode, kept as is: a failed Wolfe search leaves g_k = None and the
                    # next iteration raises (SGD:740-741, 758); the instance has terminated
                    return None
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out = ens.run(one)
            torch.cuda.synchronize()
            dt_s = time.perf_counter() - t0
            if s_ >= args.warmup:
                times.append(dt_s)
            last = (ens, out)
    finally:
        os.chdir(cwd)
    launches = lib.smo_launch_count() - n0
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dom.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ens, out = last
        crashed = sum(1 for o in out if o is None)
        out = [o for o in out if o is not None]
        its = [len(o[1]) for o in out] or [0]
        sec = float(t.item())
        line = {"metric": "SH23 optimisations/s (unmodified Optimise_On_Multi_Sphere, batched CUDA callables)", "value": total / sec, "unit": "optimisations/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak" if False else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: %d independent SH23 optimisations (Npts=256, dt=0.1, N_ITERS=500, M_0 in [0.05,0.1], LS_wolfe + CG, alpha_k=pi, "
                                       "max_iters=%d), %d per GPU, one optimiser thread per instance, calls served by batched launches; one step = the whole ensemble"
                                       % (total, iters, nb), "instances": total, "max_iters": iters},
                "ensemble": {"iterations_min_mean_max": [int(min(its)), float(np.mean(its)), int(max(its))],
                             "ended_by_the_reference_line_search_failure": crashed,
                             "batched_rounds": ens.rounds, "calls_served": ens.served,
                             "mean_group_size": {k: (ens.served[k] / max(ens.rounds[k], 1)) for k in ens.rounds},
                             "J_final_first_last": ([float(out[0][1][-1]), float(out[-1][1][-1])] if out else None)},
                "gpu_launches": int(launches),
                "e2e": {"value": total / sec, "unit": "optimisations/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                        "note": "host vectors in and out of every call (Mode H); the figure above IS end to end"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_gpu_vec(args):
    """rows A5/B5, C1-C3: Inner_Product, axpby, tangent/transport projection, retraction on dynamo-sized device vectors"""
    torch, dist, world, rank, local = _init_dist()
    from spheremanopt_b200 import _cabi
    lib = _cabi.load()
    n = 3 * 192 ** 3
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    out = torch.empty_like(x)
    work = torch.empty(lib.smo_vec_work_bytes(n), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ops = {
        # name: (call, algorithmic bytes per call: SURVEY 8(d))
        "inner_product": (lambda: lib.smo_vec_dot_dev(x.data_ptr(), y.data_ptr(), n, 1.0 / 192 ** 3, work.data_ptr(), st), 2 * n * 8),
        "axpby": (lambda: lib.smo_vec_axpby(0.3, x.data_ptr(), -2.0, y.data_ptr(), out.data_ptr(), n, st), 3 * n * 8),
        "tangent_project": (lambda: lib.smo_vec_project(x.data_ptr(), y.data_ptr(), out.data_ptr(), n, work.data_ptr(), st), 5 * n * 8),
        "retract": (lambda: lib.smo_vec_retract(x.data_ptr(), 0.7, y.data_ptr(), 1.0, 1.0 / 192 ** 3, out.data_ptr(), n, work.data_ptr(), st), 5 * n * 8),
    }
    peak, peak_src = peaks()
    res = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.smo_launch_count()
    tot_ms, tot_bytes = 0.0, 0.0
    for name, (call, nbytes) in ops.items():
        for _ in range(max(args.warmup, 3)):
            rc = call()
            assert rc == 0, lib.smo_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(args.steps, 1) * 10
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": ms, "GB/s": nbytes / (ms * 1e-3) / 1e9, "frac_of_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": nbytes}
        tot_ms += ms; tot_bytes += nbytes
    launches = lib.smo_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: Inner_Product of two pinned host vectors through the host-facing call (copies inside)
    xh = torch.empty(n, dtype=torch.float64, pin_memory=True); xh.copy_(x)
    yh = torch.empty(n, dtype=torch.float64, pin_memory=True); yh.copy_(y)
    outv = C.c_double()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        xd = xh.to("cuda", non_blocking=True); yd = yh.to("cuda", non_blocking=True)
        lib.smo_vec_dot(xd.data_ptr(), yd.data_ptr(), n, 1.0 / 192 ** 3, C.byref(outv), work.data_ptr(), st)
    t1 = time.perf_counter()
    if rank == 0:
        ip = res["inner_product"]
        line = {"metric": "vector kernels GB/s (Inner_Product, axpby, tangent projection, retraction)", "value": tot_bytes / (tot_ms * 1e-3) / 1e9 * world,
                "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "rows A5/B5, C1-C3 on device vectors of 3*192^3 doubles (170 MB each, larger than the 126 MB L2); one step = one call of each of the 4 operations, 10 repetitions each per step"},
                "kernels": res, "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": 2 * n * 8 / ((t1 - t0) / max(args.steps, 1)) / 1e9, "unit": "GB/s (Inner_Product of two pinned host vectors, H2D inside)",
                        "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": 8},
                "roofline": {"bound": "hbm", "kernel": "VecKernel<V_DOT> + FinalSum (Inner_Product)", "achieved": ip["GB/s"], "peak": peak, "unit": "GB/s",
                             "frac": ip["frac_of_peak"], "traffic": None, "launch_ms": ip["ms"], "algorithmic_bytes_per_launch": ip["algorithmic_bytes"],
                             "peak_source": peak_src}}
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
    elif args.workload == "sh23opt":
        run_gpu_sh23opt(args)
    elif args.workload == "vec":
        run_gpu_vec(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
