"""Generate the trajectory-replay fixtures (run in the build container, where /root/reference is mounted).

    python tests/golden/make_trajectories.py

The UNMODIFIED reference optimiser (``Sphere_Grad_Descent.Optimise_On_Multi_Sphere``, imported from /root/reference) is run
on the oracle's callables and EVERY point at which it evaluates ``f`` or ``Grad_f`` is recorded together with the value /
a digest of the gradient.  The GPU tests evaluate the CUDA callables at exactly those points (tests/test_gpu_history.py):
parity along a real optimisation path - line-search trial points, CG directions, retractions - without needing the
optimiser on the GPU box.  Also written: the full-length BASELINE config 2 case (24^3, 1000 steps).

 * trajectory_sh23_config1.npz - SH23 config 1 (Npts=256, dt=0.1, N_ITERS=500, E_0=0.0725), 25 iterations of SH:783's call;
                                 X of every call stored in full (512 doubles each).
 * trajectory_kdyn_N16.npz     - dynamo Npts=16, Rm=1, dt=1e-3, N_ITERS=40, 4 iterations of KD:1066's call; the iterates
                                 are band limited (Generate_IC output + band-limited gradients), so X is stored as its
                                 retained Fourier coefficients (8x15x15 complex per component) and rebuilt with the
                                 oracle's to_grid in the test.
 * config2_kdyn24.json         - f, Grad_f digests of the oracle at BASELINE config 2 (Npts=24, Rm=1, dt=1e-3, N_ITERS=1000,
                                 Generate_IC(Noise=True) inputs).
Parity unpinned against Dedalus (oracle/__init__.py): these fixtures pin the oracle + the real reference optimiser.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "refstubs"))
sys.path.insert(0, "/root/reference")

from oracle import kdyn as okd   # noqa: E402
from oracle import sh23 as osh   # noqa: E402
from tests.golden.make_golden import digest   # noqa: E402

import Sphere_Grad_Descent as SGD   # noqa: E402  (unmodified reference)


def gdigest(v):
    d = digest(v)
    return np.array([d["sum"], d["sumsq"], d["wdot"]] + d["sample"][:64], dtype=np.float64)


class Recorder:
    def __init__(self, f, g, pack):
        self.f, self.g, self.pack = f, g, pack
        self.kind, self.X, self.val, self.gd = [], [], [], []

    def fwd(self, X, *a):
        v = self.f(X, *a)
        self.kind.append(0); self.X.append(self.pack(X)); self.val.append(float(v)); self.gd.append(None)
        return v

    def adj(self, X, *a):
        g = self.g(X, *a)
        self.kind.append(1); self.X.append(self.pack(X)); self.val.append(np.nan); self.gd.append([gdigest(gi) for gi in g])
        return g

    def save(self, path, **meta):
        ng = max(len(g) for g in self.gd if g is not None)
        gd = np.full((len(self.kind), ng, 67), np.nan)
        for i, g in enumerate(self.gd):
            if g is not None:
                gd[i] = np.stack(g)
        np.savez_compressed(path, kind=np.array(self.kind, dtype=np.int8), X=np.stack(self.X), val=np.array(self.val), gdigest=gd,
                            **{k: np.asarray(v) for k, v in meta.items()})
        print("wrote %s: %d calls (%d f, %d Grad_f), %.1f KB" % (path, len(self.kind), self.kind.count(0), self.kind.count(1),
                                                                os.path.getsize(path) / 1e3))


def main():
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())   # the reference writes optimize_result.txt etc. into the CWD
    try:
        # ---- SH23 config 1 ------------------------------------------------------------------------------------------
        E_0, nit, dt = 0.0725, 500, 0.1
        dom, X0 = osh.Generate_IC(E_0)
        D = osh.GEN_BUFFER(dom, nit)
        rec = Recorder(osh.FWD_Solve_IVP_Lin, osh.ADJ_Solve_IVP_Lin, lambda X: np.array(X[0], dtype=np.float64))
        args_f = [dom, dt, nit, nit, D, None, "Discrete"]
        RES, FUN, Xopt = SGD.Optimise_On_Multi_Sphere([X0], [E_0], rec.fwd, rec.adj, osh.Inner_Prod, args_f, (dom, None), max_iters=25,
                                                      alpha_k=np.pi, LS='LS_wolfe', CG=True, callback=None, verbose=False)
        rec.save(os.path.join(HERE, "trajectory_sh23_config1.npz"), RESIDUAL=np.array(RES, dtype=np.float64), FUNCT=np.array(FUN, dtype=np.float64),
                 dt=dt, N_ITERS=nit, E_0=E_0)
        # ---- dynamo Npts = 16 ---------------------------------------------------------------------------------------
        Npts, nit = 16, 40
        dom, B0, U = okd.Generate_IC(Npts, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3)

        def pack(X):
            out = []
            for v in X:
                comps = okd.Vec_to_Field(dom, v)
                c = [dom.to_coef_3d(ci) for ci in comps]
                back = okd.Field_to_Vec(dom, *[dom.to_grid_3d(ci) for ci in c])
                assert np.abs(back - v).max() <= 1e-13 * np.abs(v).max(), "iterate is not band limited"
                out.append(np.stack(c))
            return np.stack(out)     # [2][3][8][15][15] complex
        D = okd.GEN_BUFFER(Npts, dom, nit)
        rec = Recorder(okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, pack)
        args_f = [dom, 1.0, 1e-3, nit, nit, D, "Final", "Discrete"]
        RES, FUN, Xopt = SGD.Optimise_On_Multi_Sphere([B0, U], [1.0, 1.0], rec.fwd, rec.adj, okd.Inner_Prod_3, args_f, (dom, None), max_iters=4,
                                                      alpha_k=100., LS='LS_wolfe', CG=True, callback=None, verbose=False)
        rec.save(os.path.join(HERE, "trajectory_kdyn_N16.npz"), RESIDUAL=np.array(RES, dtype=np.float64), FUNCT=np.array(FUN, dtype=np.float64),
                 Rm=1.0, dt=1e-3, N_ITERS=nit, Npts=Npts)
        # ---- BASELINE config 2 at full length -----------------------------------------------------------------------
        Npts, nit, Rm, dt = 24, 1000, 1.0, 1e-3
        dom, B0, U = okd.Generate_IC(Npts, (0., 2. * np.pi), 1.0, True, Rm=Rm, dt=dt)
        D = okd.GEN_BUFFER(Npts, dom, nit)
        t0 = time.time()
        f = okd.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, D)
        g = okd.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, D)
        out = {"Npts": Npts, "Rm": Rm, "dt": dt, "N_ITERS": nit, "ic": "Generate_IC(24, (0,2pi), 1.0, True, Rm=1.0, dt=1e-3)", "B0": digest(B0), "U": digest(U),
               "f": float(f), "gradB": digest(g[0]), "gradU": digest(g[1]), "oracle_seconds": time.time() - t0}
        with open(os.path.join(HERE, "config2_kdyn24.json"), "w") as fh:
            json.dump(out, fh, indent=1)
        print("wrote config2_kdyn24.json (oracle %.1f s)" % out["oracle_seconds"])
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
