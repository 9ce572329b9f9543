"""CPU tests of the oracle (numpy restatement of the reference algorithm): committed golden vectors, the reference's
own acceptance test (Taylor remainder, run with the UNMODIFIED TestGrad.py when /root/reference is mounted), closed-form
linear decay of both time steppers and the adjoint dot-product identity."""
import json
import os

import numpy as np
import pytest

from oracle import kdyn as okd
from oracle import sh23 as osh
from tests.common import kdyn_field, sh23_input

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
TOL = 1e-11


def digest_close(v, d, tol=TOL):
    v = np.asarray(v, dtype=np.float64).ravel()
    assert v.size == d["n"]
    w = np.cos(0.37 * np.arange(v.size) + 0.1)
    scale = np.sqrt(d["sumsq"]) + 1e-300
    assert abs(v.sum() - d["sum"]) <= tol * scale * np.sqrt(v.size)
    assert abs((v * v).sum() - d["sumsq"]) <= tol * d["sumsq"] + 1e-300
    assert abs((v * w).sum() - d["wdot"]) <= tol * scale * np.sqrt(v.size)
    s = np.asarray(d["sample"])
    assert np.abs(v[::d["sample_stride"]] - s).max() <= tol * max(np.abs(s).max(), 1e-300) * 10


@pytest.mark.parametrize("Npts", [64, 128, 256])
def test_sh23_golden(Npts):
    g = GOLD["case_sh23_N%d" % Npts]
    dom = osh.domain_sh23(Npts)
    X = sh23_input(dom, seed=Npts)
    digest_close(X, g["X"])
    D = osh.GEN_BUFFER(dom, g["N_ITERS"])
    f = osh.FWD_Solve_IVP_Lin([X], dom, g["dt"], g["N_ITERS"], g["N_ITERS"], D)
    assert abs(f - g["f"]) <= TOL * abs(g["f"])
    digest_close(osh.ADJ_Solve_IVP_Lin([X], dom, g["dt"], g["N_ITERS"], g["N_ITERS"], D)[0], g["grad"])
    digest_close(osh.ADJ_Solve_IVP_Lin([X], dom, g["dt"], g["N_ITERS"], g["N_ITERS"], D, None, "Continuous")[0], g["grad_continuous"])
    digest_close(D['A_fwd'][:, -1].real, g["snap_last_re"])


@pytest.mark.parametrize("Npts", [16, 24])
def test_kdyn_golden(Npts):
    g = GOLD["case_kdyn_N%d" % Npts]
    dom = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(dom, 1), kdyn_field(dom, 2)
    digest_close(B0, g["B0"])
    nit = g["N_ITERS"]
    D = okd.GEN_BUFFER(Npts, dom, nit)
    f = okd.FWD_Solve_IVP_Lin([B0, U], dom, g["Rm"], g["dt"], nit, nit, D)
    assert abs(f - g["f"]) <= TOL * abs(g["f"])
    gr = okd.ADJ_Solve_IVP_Lin([B0, U], dom, g["Rm"], g["dt"], nit, nit, D)
    digest_close(gr[0], g["gradB"])
    digest_close(gr[1], g["gradU"])


def test_initial_conditions_golden():
    _, X0 = osh.Generate_IC(0.0725)
    digest_close(X0, GOLD["ic_sh23_config1"])
    _, B, U = okd.Generate_IC(16, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3)
    digest_close(B, GOLD["ic_kdyn_N16"]["B"])
    digest_close(U, GOLD["ic_kdyn_N16"]["U"])


def test_kdyn_integrated_golden():
    g = GOLD["case_kdyn_N16_integrated"]
    dom, B0, U = okd.Generate_IC(16, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3)
    nit = g["N_ITERS"]
    D = okd.GEN_BUFFER(16, dom, nit)
    args = [dom, 1.0, 1e-3, nit, nit, D, "Integrated", "Discrete"]
    f = okd.FWD_Solve_IVP_Lin([B0, U], *args)
    assert abs(f - g["f"]) <= TOL * abs(g["f"])
    gr = okd.ADJ_Solve_IVP_Lin([B0, U], *args)
    digest_close(gr[0], g["gradB"])
    digest_close(gr[1], g["gradU"])


def test_golden_taylor_slopes_are_two():
    """the table saved by the unmodified reference Adjoint_Gradient_Test (make_golden.py): R ~ h, R2 ~ h^2"""
    for key in ("taylor_sh23", "taylor_kdyn_dB", "taylor_kdyn_dBdU", "taylor_kdyn_integrated"):
        AA = np.asarray(GOLD[key])
        assert np.all(np.abs(AA[3, :4] - 1.0) < 2e-2), key
        assert np.all(np.abs(AA[4, :4] - 2.0) < 1e-2), key


def test_reference_gradient_test_on_oracle(refopt, tmp_path, monkeypatch):
    """run the UNMODIFIED reference TestGrad.py on the oracle callables (only where /root/reference is mounted)"""
    SGD, TG = refopt
    monkeypatch.chdir(tmp_path)
    dom, X0 = osh.Generate_IC(1.0)
    nit = 100
    D = osh.GEN_BUFFER(dom, nit)
    TG.Adjoint_Gradient_Test(X0, X0.copy(), osh.FWD_Solve_IVP_Lin, osh.ADJ_Solve_IVP_Lin, osh.Inner_Prod,
                             [dom, 0.1, nit, nit, D, None, "Discrete"], (dom, None), epsilon=1e-4)
    AA = np.load("eps_TestR_TestR2_h_h2.npy")
    assert np.all(np.abs(AA[4, :4] - 2.0) < 1e-2)


def test_reference_optimiser_history_matches_golden(refopt, tmp_path, monkeypatch):
    """the UNMODIFIED Optimise_On_Multi_Sphere on the oracle reproduces the committed RESIDUAL / FUNCT history"""
    SGD, TG = refopt
    monkeypatch.chdir(tmp_path)
    g = GOLD["history_kdyn_N16"]
    nit = g["N_ITERS"]
    dom, B0, U = okd.Generate_IC(16, (0., 2. * np.pi), 1.0, True, Rm=1.0, dt=1e-3)
    D = okd.GEN_BUFFER(16, dom, nit)
    RES, FUN, _ = SGD.Optimise_On_Multi_Sphere([B0, U], [1.0, 1.0], okd.FWD_Solve_IVP_Lin, okd.ADJ_Solve_IVP_Lin, okd.Inner_Prod_3,
                                               [dom, 1.0, 1e-3, nit, nit, D, "Final", "Discrete"], (dom, None), max_iters=g["max_iters"],
                                               alpha_k=100., LS='LS_wolfe', CG=True, callback=None, verbose=False)
    assert np.allclose(FUN, g["FUNCT"], rtol=1e-10, atol=0)
    assert np.allclose(np.asarray(RES, dtype=float), np.asarray(g["RESIDUAL"]), rtol=1e-8, atol=0)


def test_sh23_linear_decay_closed_form():
    """tiny amplitude => nonlinearity negligible: every mode decays by 1/(1 + dt L_k) per SBDF1 step"""
    dom = osh.domain_sh23(64)
    c = np.zeros(dom.Nh, dtype=complex); c[5] = 1e-14 * (1 + 2j); c[11] = 2e-14j
    X = dom.to_grid_1d(c)
    nit, dt = 7, 0.1
    D = osh.GEN_BUFFER(dom, nit)
    osh.FWD_Solve_IVP_Lin([X], dom, dt, nit, nit, D)
    Lk = osh._Lk(dom)
    want = c / (1.0 + dt * Lk) ** nit
    assert np.abs(D["A_fwd"][:, -1] - want).max() <= 1e-11 * np.abs(want).max()


def test_kdyn_linear_decay_closed_form():
    """U = 0: CNAB1 diffusion of a solenoidal mode, factor (1/dt - k^2/2Rm)/(1/dt + k^2/2Rm) per step"""
    dom = okd.domain_kdyn(16)
    B0 = kdyn_field(dom, 5)
    U = np.zeros_like(B0)
    nit, dt, Rm = 5, 1e-2, 3.0
    D = okd.GEN_BUFFER(16, dom, nit)
    okd.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, D)
    K2 = okd._K(dom)[3]
    fac = ((1 / dt - K2 / (2 * Rm)) / (1 / dt + K2 / (2 * Rm))) ** nit
    for key in ('A_fwd', 'B_fwd', 'C_fwd'):
        want = D[key][..., 0] * fac
        assert np.abs(D[key][..., -1] - want).max() <= 1e-12 * np.abs(want).max()


def test_adjoint_dot_product_identity():
    """<Grad_f, dX> equals the directional derivative of f to O(eps^2) (central difference), both examples"""
    dom = okd.domain_kdyn(16)
    B0, U, dB, dU = (kdyn_field(dom, s) for s in (1, 2, 3, 4))
    nit = 8
    D = okd.GEN_BUFFER(16, dom, nit)
    a = [dom, 1.0, 1e-3, nit, nit, D]
    okd.FWD_Solve_IVP_Lin([B0, U], *a)
    g = okd.ADJ_Solve_IVP_Lin([B0, U], *a)
    lhs = okd.Inner_Prod_3(g[0], dB, dom) + okd.Inner_Prod_3(g[1], dU, dom)
    e = 1e-5
    rhs = (okd.FWD_Solve_IVP_Lin([B0 + e * dB, U + e * dU], *a) - okd.FWD_Solve_IVP_Lin([B0 - e * dB, U - e * dU], *a)) / (2 * e)
    assert abs(lhs - rhs) <= 1e-7 * abs(rhs)
