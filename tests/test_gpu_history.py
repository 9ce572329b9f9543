"""GPU parity along real optimisation paths (BASELINE north_star: "the optimiser's RESIDUAL/FUNCT history to 1e-9").

 * trajectory replay: the CUDA callables are evaluated at EVERY point at which the unmodified reference optimiser evaluated
   the oracle's callables (fixtures of tests/golden/make_trajectories.py): f to 1e-9, Grad_f digests to 1e-9;
 * the unmodified ``Optimise_On_Multi_Sphere`` (SGD:692) and ``Adjoint_Gradient_Test`` (TG:5) themselves, imported from
   $SMO_REFERENCE_DIR | /root/reference | baseline/_ref, driven by the CUDA callables: RESIDUAL / FUNCT histories and the
   Taylor table against the golden ones;
 * BASELINE config 2 at its full length (24^3, 1000 steps each way) against the oracle's digests.
"""
import json
import os
import tempfile
import time

import numpy as np
import pytest

from oracle import kdyn as okd
from oracle import sh23 as osh

pytestmark = pytest.mark.gpu
TOL = 1e-9
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest_vec(v):
    v = np.asarray(v, dtype=np.float64).ravel()
    w = np.cos(0.37 * np.arange(v.size) + 0.1)
    stride = max(1, v.size // 64)
    return np.concatenate([[v.sum(), (v * v).sum(), (v * w).sum()], v[::stride][:64]])


def _digest_close(got, want, n):
    """digest = [sum, sumsq, weighted sum, 64 strided samples]; sums may cancel, so their error is measured against the
    vector's scale sqrt(n * sumsq)"""
    scale = np.sqrt(n * want[1])
    assert abs(got[1] - want[1]) <= TOL * want[1]
    assert abs(got[0] - want[0]) <= TOL * scale and abs(got[2] - want[2]) <= TOL * scale
    ns = min(len(got), len(want)) - 3
    assert np.abs(got[3:3 + ns] - want[3:3 + ns]).max() <= TOL * np.abs(want[3:3 + ns]).max()


def _dict_digest_close(v, d):
    got = _digest_vec(v)
    want = np.array([d["sum"], d["sumsq"], d["wdot"]] + d["sample"][:64])
    _digest_close(got, want, d["n"])


def test_trajectory_replay_sh23_config1():
    """every f / Grad_f evaluation of 25 optimiser iterations of BASELINE config 1, replayed on the CUDA callables"""
    from spheremanopt_b200 import sh23
    T = np.load(os.path.join(GOLD, "trajectory_sh23_config1.npz"))
    dt, nit = float(T["dt"]), int(T["N_ITERS"])
    dom = sh23.Domain(256)
    store = sh23.GEN_BUFFER(dom, nit)
    args = (dom, dt, nit, nit, store, None, "Discrete")
    worst_f = 0.0
    for kind, X, val, gd in zip(T["kind"], T["X"], T["val"], T["gdigest"]):
        f = sh23.FWD_Solve_IVP_Lin([X], *args)      # (Grad_f replays the store the preceding f(X) wrote, SGD:740-796)
        if kind == 0:
            worst_f = max(worst_f, abs(f - val) / abs(val))
            assert abs(f - val) <= TOL * abs(val)
        else:
            g = sh23.ADJ_Solve_IVP_Lin([X], *args)
            _digest_close(_digest_vec(g[0]), gd[0], X.size)
    assert int((T["kind"] == 0).sum()) >= 25 and int((T["kind"] == 1).sum()) >= 10
    print("SH23 trajectory replay: %d calls, worst rel.err of f %.2e" % (len(T["kind"]), worst_f))


def test_trajectory_replay_kdyn_N16():
    from spheremanopt_b200 import kdyn
    T = np.load(os.path.join(GOLD, "trajectory_kdyn_N16.npz"))
    Npts, nit, Rm, dt = int(T["Npts"]), int(T["N_ITERS"]), float(T["Rm"]), float(T["dt"])
    od = okd.domain_kdyn(Npts)
    dom = kdyn.Domain(Npts)
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    args = (dom, Rm, dt, nit, nit, store, "Final", "Discrete")
    for kind, Xc, val, gd in zip(T["kind"], T["X"], T["val"], T["gdigest"]):
        X = [okd.Field_to_Vec(od, *[od.to_grid_3d(c) for c in Xc[i]]) for i in range(2)]     # iterates are band limited
        f = kdyn.FWD_Solve_IVP_Lin(X, *args)
        if kind == 0:
            assert abs(f - val) <= TOL * abs(val)
        else:
            g = kdyn.ADJ_Solve_IVP_Lin(X, *args)
            _digest_close(_digest_vec(g[0]), gd[0], X[0].size)
            _digest_close(_digest_vec(g[1]), gd[1], X[1].size)


def test_config2_full_length():
    """BASELINE config 2: 24^3, Rm=1, dt=1e-3, N_ITERS=1000, Generate_IC(Noise=True) inputs - f and Grad_f against the
    oracle's digests, host vectors (Mode H) through the reference-facing callables; also prints the latency-bound rate"""
    from spheremanopt_b200 import kdyn
    G = json.load(open(os.path.join(GOLD, "config2_kdyn24.json")))
    Npts, nit, Rm, dt = G["Npts"], G["N_ITERS"], G["Rm"], G["dt"]
    dom, B0, U = kdyn.Generate_IC(Npts, (0., 2. * np.pi), 1.0, True, Rm=Rm, dt=dt)
    _dict_digest_close(B0, G["B0"]); _dict_digest_close(U, G["U"])          # the CUDA Generate_IC reproduces the oracle's inputs
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    args = (dom, Rm, dt, nit, nit, store, "Final", "Discrete")
    for rep in range(2):
        t0 = time.perf_counter()
        f = kdyn.FWD_Solve_IVP_Lin([B0, U], *args)
        g = kdyn.ADJ_Solve_IVP_Lin([B0, U], *args)
        t1 = time.perf_counter()
    assert abs(f - G["f"]) <= TOL * abs(G["f"])
    _dict_digest_close(g[0], G["gradB"]); _dict_digest_close(g[1], G["gradU"])
    print("config 2 (24^3 x 1000 steps): %.1f ms per Grad_f pair = %.0f time steps/s (oracle: %.1f s)" % ((t1 - t0) * 1e3, 2 * nit / (t1 - t0), G["oracle_seconds"]))


def _run_in_tmp(fn):
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())    # the reference writes optimize_result.txt / eps_*.npy / DAL_PROGRESS.h5 into the CWD
    try:
        return fn()
    finally:
        os.chdir(cwd)


def _hist_close(got, want, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want) / np.abs(want)
    print("%s: %d entries, worst rel.err %.2e" % (what, got.size, err.max()))
    assert err.max() <= TOL, (what, err)


def test_reference_optimiser_history_sh23(refopt):
    """the UNMODIFIED Optimise_On_Multi_Sphere on the CUDA callables, BASELINE config 1 (SH:783), 25 iterations"""
    SGD, _ = refopt
    from spheremanopt_b200 import sh23
    gold = json.load(open(os.path.join(GOLD, "golden.json")))["history_sh23_config1"]
    E_0, nit = 0.0725, 500
    dom, X0 = sh23.Generate_IC(E_0)
    store = sh23.GEN_BUFFER(dom, nit)
    args_f = [dom, 0.1, nit, nit, store, None, "Discrete"]
    RES, FUN, Xopt = _run_in_tmp(lambda: SGD.Optimise_On_Multi_Sphere([X0], [E_0], sh23.FWD_Solve_IVP_Lin, sh23.ADJ_Solve_IVP_Lin, sh23.Inner_Prod,
                                                                      args_f, (dom, None), max_iters=gold["max_iters"], alpha_k=np.pi, LS='LS_wolfe',
                                                                      CG=True, callback=None, verbose=False))
    _hist_close(FUN, gold["FUNCT"], "SH23 FUNCT")
    _hist_close(RES, gold["RESIDUAL"], "SH23 RESIDUAL")
    _dict_digest_close(Xopt[0], gold["X_opt"])


@pytest.mark.parametrize("mode", ["H", "D"])
def test_reference_optimiser_history_kdyn(refopt, mode):
    """the same for the dynamo (KD:1066's call at Npts=16): host vectors (Mode H) and device-resident DevVec (Mode D)"""
    SGD, _ = refopt
    from spheremanopt_b200 import kdyn
    from spheremanopt_b200.devvec import DevVec
    gold = json.load(open(os.path.join(GOLD, "golden.json")))["history_kdyn_N16"]
    Npts, nit = 16, gold["N_ITERS"]
    dom, B0, U = kdyn.Generate_IC(Npts, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3, as_devvec=(mode == "D"))
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    args_f = [dom, 1.0, 1e-3, nit, nit, store, "Final", "Discrete"]
    RES, FUN, Xopt = _run_in_tmp(lambda: SGD.Optimise_On_Multi_Sphere([B0, U], [1.0, 1.0], kdyn.FWD_Solve_IVP_Lin, kdyn.ADJ_Solve_IVP_Lin,
                                                                      kdyn.Inner_Prod_3, args_f, (dom, None), max_iters=gold["max_iters"],
                                                                      alpha_k=100., LS='LS_wolfe', CG=True, callback=None, verbose=False))
    _hist_close(FUN, gold["FUNCT"], "KDyn FUNCT (mode %s)" % mode)
    _hist_close(RES, gold["RESIDUAL"], "KDyn RESIDUAL (mode %s)" % mode)
    Bopt = Xopt[0].numpy() if isinstance(Xopt[0], DevVec) else Xopt[0]
    _dict_digest_close(Bopt, gold["B_opt"])


def test_reference_adjoint_gradient_test_sh23(refopt):
    """the reference's own acceptance test (TG:5) on the CUDA callables, as shipped at SH:773-778 (X_0 = dX_0 = Generate_IC(1.))"""
    _, TG = refopt
    from spheremanopt_b200 import sh23
    gold = np.array(json.load(open(os.path.join(GOLD, "golden.json")))["taylor_sh23"])
    dom, X0 = sh23.Generate_IC(1.0)
    nit = 500
    store = sh23.GEN_BUFFER(dom, nit)
    args_f = [dom, 0.1, nit, nit, store, None, "Discrete"]

    def run():
        TG.Adjoint_Gradient_Test(X0, X0.copy(), sh23.FWD_Solve_IVP_Lin, sh23.ADJ_Solve_IVP_Lin, sh23.Inner_Prod, args_f, (dom, None), epsilon=1e-4)
        return np.load("eps_TestR_TestR2_h_h2.npy")
    tab = _run_in_tmp(run)          # rows: eps, R (1st order), R2 (2nd order), slopes of R, slopes of R2 (TG:129-154)
    assert tab.shape == gold.shape
    assert np.abs(tab[1] - gold[1]).max() <= 1e-6 * np.abs(gold[1]).max()          # first-order remainders
    k = min(4, tab.shape[1])
    print("Taylor slopes R2 (CUDA callables):", np.round(tab[4][:k], 4), " golden:", np.round(gold[4][:k], 4))
    assert np.all(np.abs(tab[4][:k] - 2.0) < 0.05) and np.all(np.abs(tab[4][:k] - gold[4][:k]) < 0.02)
