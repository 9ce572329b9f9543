"""Generate the committed golden fixtures (run in the build container, where /root/reference is mounted).

    python tests/golden/make_golden.py

What is pinned, and by what:
 * ``taylor_*``   - the UNMODIFIED reference ``TestGrad.Adjoint_Gradient_Test`` (imported from /root/reference) run on
                    the oracle's callables: the Taylor-remainder table it saves (eps, R, R2, slopes).  This is the
                    reference's own acceptance test for the path.
 * ``history_*``  - the UNMODIFIED reference ``Sphere_Grad_Descent.Optimise_On_Multi_Sphere`` run on the oracle's
                    callables: RESIDUAL / FUNCT histories (the north-star's history-parity quantity).
 * ``case_*``     - oracle outputs (J, gradients, snapshot samples) on seeded inputs, so that a later change of the
                    oracle itself is caught; the GPU tests compare the CUDA path with the same numbers.
The arithmetic below the reference's callable boundary lives in Dedalus, which cannot be installed here, so the
fixtures pin the oracle + the real reference optimiser, not Dedalus (see oracle/__init__.py: PARITY UNPINNED).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "refstubs"))
sys.path.insert(0, "/root/reference")

from oracle import kdyn as okd   # noqa: E402
from oracle import sh23 as osh   # noqa: E402
from tests.common import kdyn_field, sh23_input   # noqa: E402

import Sphere_Grad_Descent as SGD   # noqa: E402  (unmodified reference)
import TestGrad as TG               # noqa: E402  (unmodified reference)


def digest(v):
    """order-sensitive fingerprints of a vector + a strided sample"""
    v = np.asarray(v, dtype=np.float64).ravel()
    w = np.cos(0.37 * np.arange(v.size) + 0.1)
    return {"n": int(v.size), "sum": float(v.sum()), "sumsq": float((v * v).sum()), "wdot": float((v * w).sum()),
            "sample_stride": max(1, v.size // 64), "sample": v[::max(1, v.size // 64)].tolist()}


def taylor(X0, dX0, f, g, ip, args_f, args_IP):
    TG.Adjoint_Gradient_Test(X0, dX0, f, g, ip, args_f, args_IP, epsilon=1e-4)
    return np.load("eps_TestR_TestR2_h_h2.npy").tolist()


def main():
    out = {}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)   # the reference writes optimize_result.txt / eps_*.npy into the CWD
    try:
        # ---- SH23 cases ------------------------------------------------------------------------------------
        for Npts, dt, nit in [(64, 0.1, 30), (128, 0.05, 40), (256, 0.1, 500)]:
            dom = osh.domain_sh23(Npts)
            X = sh23_input(dom, seed=Npts)
            D = osh.GEN_BUFFER(dom, nit)
            f = osh.FWD_Solve_IVP_Lin([X], dom, dt, nit, nit, D)
            g = osh.ADJ_Solve_IVP_Lin([X], dom, dt, nit, nit, D)[0]
            gc = osh.ADJ_Solve_IVP_Lin([X], dom, dt, nit, nit, D, None, "Continuous")[0]
            out["case_sh23_N%d" % Npts] = {"Npts": Npts, "dt": dt, "N_ITERS": nit, "X": digest(X), "f": f,
                                           "grad": digest(g), "grad_continuous": digest(gc),
                                           "snap_last_re": digest(D['A_fwd'][:, -1].real),
                                           "snap_last_im": digest(D['A_fwd'][:, -1].imag)}
        # ---- SH23: the reference's shipped gradient test (SH:773-778: X_0 = dX_0 = Generate_IC(1.)), config-1 params
        dom, X0 = osh.Generate_IC(1.0)
        nit = 500
        D = osh.GEN_BUFFER(dom, nit)
        args_f = [dom, 0.1, nit, nit, D, None, "Discrete"]
        out["taylor_sh23"] = taylor(X0, X0.copy(), osh.FWD_Solve_IVP_Lin, osh.ADJ_Solve_IVP_Lin, osh.Inner_Prod, args_f, (dom, None))
        out["ic_sh23_E1"] = digest(X0)
        # ---- SH23 config 1: optimiser history (first 25 iterations of the 200 of SH:783)
        E_0 = 0.0725
        dom, X0 = osh.Generate_IC(E_0)
        out["ic_sh23_config1"] = digest(X0)
        D = osh.GEN_BUFFER(dom, nit)
        args_f = [dom, 0.1, nit, nit, D, None, "Discrete"]
        RES, FUN, Xopt = SGD.Optimise_On_Multi_Sphere([X0], [E_0], osh.FWD_Solve_IVP_Lin, osh.ADJ_Solve_IVP_Lin, osh.Inner_Prod,
                                                      args_f, (dom, None), max_iters=25, alpha_k=np.pi, LS='LS_wolfe', CG=True,
                                                      callback=None, verbose=False)
        out["history_sh23_config1"] = {"max_iters": 25, "RESIDUAL": [list(map(float, r)) for r in RES],
                                       "FUNCT": list(map(float, FUN)), "X_opt": digest(Xopt[0])}
        # ---- KDyn cases ------------------------------------------------------------------------------------
        for Npts, nit in [(16, 6), (24, 25)]:
            dom = okd.domain_kdyn(Npts)
            B0, U = kdyn_field(dom, 1), kdyn_field(dom, 2)
            D = okd.GEN_BUFFER(Npts, dom, nit)
            f = okd.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, D)
            g = okd.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, D)
            out["case_kdyn_N%d" % Npts] = {"Npts": Npts, "Rm": 1.0, "dt": 1e-3, "N_ITERS": nit, "B0": digest(B0), "U": digest(U),
                                           "f": f, "gradB": digest(g[0]), "gradU": digest(g[1]),
                                           "snapA_last_re": digest(D['A_fwd'][..., -1].real)}
        # ---- KDyn gradient test (KD:1057-1060: dX = [dB, 0*dU]) and one perturbing U as well, Npts = 16
        Npts, nit = 16, 40
        dom, B0, U = okd.Generate_IC(Npts, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3)
        out["ic_kdyn_N16"] = {"B": digest(B0), "U": digest(U)}
        D = okd.GEN_BUFFER(Npts, dom, nit)
        args_f = [dom, 1.0, 1e-3, nit, nit, D, "Final", "Discrete"]
        dB, dU = kdyn_field(dom, 11), kdyn_field(dom, 12)
        out["taylor_kdyn_dB"] = taylor([B0, U], [dB, 0. * dU], okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, okd.Inner_Prod_3, args_f, (dom, None))
        out["taylor_kdyn_dBdU"] = taylor([B0, U], [dB, dU], okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, okd.Inner_Prod_3, args_f, (dom, None))
        # ---- KDyn, Cost_function="Integrated" (KD:655-669, 738-742, 861-864): the reference's gradient test + digests
        DI = okd.GEN_BUFFER(Npts, dom, nit)
        args_i = [dom, 1.0, 1e-3, nit, nit, DI, "Integrated", "Discrete"]
        out["taylor_kdyn_integrated"] = taylor([B0, U], [dB, dU], okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, okd.Inner_Prod_3, args_i, (dom, None))
        fi = okd.FWD_Solve_IVP_Lin([B0, U], *args_i)
        gi = okd.ADJ_Solve_IVP_Lin([B0, U], *args_i)
        out["case_kdyn_N16_integrated"] = {"N_ITERS": nit, "f": fi, "gradB": digest(gi[0]), "gradU": digest(gi[1])}
        # ---- KDyn optimiser history, Npts = 16, 40 steps, 4 iterations of KD:1066's call
        RES, FUN, Xopt = SGD.Optimise_On_Multi_Sphere([B0, U], [1.0, 1.0], okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, okd.Inner_Prod_3,
                                                      args_f, (dom, None), max_iters=4, alpha_k=100., LS='LS_wolfe', CG=True,
                                                      callback=None, verbose=False)
        out["history_kdyn_N16"] = {"max_iters": 4, "N_ITERS": nit, "RESIDUAL": [list(map(float, r)) for r in RES],
                                   "FUNCT": list(map(float, FUN)), "B_opt": digest(Xopt[0]), "U_opt": digest(Xopt[1])}
    finally:
        os.chdir(cwd)
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))
    for k in ("taylor_sh23", "taylor_kdyn_dB", "taylor_kdyn_dBdU", "taylor_kdyn_integrated"):
        print(k, "slopes R:", np.round(out[k][3][:4], 4), " R2:", np.round(out[k][4][:4], 4))


if __name__ == "__main__":
    main()
