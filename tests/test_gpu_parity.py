"""GPU parity: the CUDA path, called through the reference-facing callables (which go through the C ABI of
libsmo_b200.so), against the numpy oracle on the same seeded inputs.  Tolerance: 1e-9 relative (BASELINE north_star)."""
import copy

import numpy as np
import pytest

from oracle import kdyn as okd
from oracle import sh23 as osh
from oracle import sphere as osp
from tests.common import kdyn_field, relerr, sh23_input

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("Npts,dt,nit", [(64, 0.1, 30), (128, 0.05, 40), (256, 0.1, 500)])
@pytest.mark.parametrize("adj", ["Discrete", "Continuous"])
def test_sh23_f_gradf(Npts, dt, nit, adj):
    from spheremanopt_b200 import sh23
    dom = sh23.Domain(Npts)
    od = osh.domain_sh23(Npts)
    X = sh23_input(od, seed=Npts)
    store = sh23.GEN_BUFFER(dom, nit)
    f = sh23.FWD_Solve_IVP_Lin([X], dom, dt, nit, nit, store, None, adj)
    g = sh23.ADJ_Solve_IVP_Lin([X], dom, dt, nit, nit, store, None, adj)
    D = osh.GEN_BUFFER(od, nit)
    fo = osh.FWD_Solve_IVP_Lin([X], od, dt, nit, nit, D, None, adj)
    go = osh.ADJ_Solve_IVP_Lin([X], od, dt, nit, nit, D, None, adj)
    assert abs(f - fo) <= TOL * abs(fo)
    assert relerr(store['A_fwd'], D['A_fwd']) <= TOL
    assert relerr(g[0], go[0]) <= TOL
    assert abs(sh23.Inner_Prod(X, g[0], dom) - osh.Inner_Prod(X, go[0], od)) <= TOL * abs(osh.Inner_Prod(X, go[0], od))


def test_sh23_batch_ragged():
    """batch sizes that do not fill the last CTA (4 instances per CTA) and more CTAs than one wave"""
    from spheremanopt_b200 import sh23
    dom = sh23.Domain(256)
    od = osh.domain_sh23(256)
    for batch in (1, 3, 7, 1201):
        X = np.concatenate([sh23_input(od, seed=b % 5, amp=0.03 + 0.001 * (b % 11)) for b in range(batch)])
        store = sh23.GEN_BUFFER(dom, 20, batch=batch)
        J = sh23.forward_batch(X, dom, 0.1, 20, store).cpu().numpy()
        G = sh23.adjoint_batch(dom, 0.1, 20, store).cpu().numpy().reshape(batch, -1)
        for b in sorted(set([0, batch // 2, batch - 1])):
            D = osh.GEN_BUFFER(od, 20)
            xb = X[b * od.M:(b + 1) * od.M]
            fo = osh.FWD_Solve_IVP_Lin([xb], od, 0.1, 20, 20, D)
            go = osh.ADJ_Solve_IVP_Lin([xb], od, 0.1, 20, 20, D)[0]
            assert abs(-J[b] - fo) <= TOL * abs(fo)
            assert relerr(G[b], go) <= TOL


def test_sh23_generate_ic():
    from spheremanopt_b200 import sh23
    dom, X0 = sh23.Generate_IC(0.0725)
    od, X0o = osh.Generate_IC(0.0725)
    assert relerr(X0, X0o) <= TOL
    assert abs(sh23.Inner_Prod(X0, X0, dom) - 0.0725) <= 1e-12


@pytest.mark.parametrize("Npts,nit", [(16, 6), (24, 25), (32, 5), (64, 3)])
def test_kdyn_f_gradf(Npts, nit):
    from spheremanopt_b200 import kdyn
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0 = kdyn_field(od, 1)
    U = kdyn_field(od, 2)
    Rm, dt = 1.0, 1e-3
    store = kdyn.GEN_BUFFER(Npts, dom, nit)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    assert abs(f - fo) <= TOL * abs(fo)
    for key in ('A_fwd', 'B_fwd', 'C_fwd'):
        assert relerr(store[key], D[key]) <= TOL
    assert relerr(g[0], go[0]) <= TOL
    assert relerr(g[1], go[1]) <= TOL
    ip, ipo = kdyn.Inner_Prod_3(B0, g[0], dom), okd.Inner_Prod_3(B0, go[0], od)
    assert abs(ip - ipo) <= TOL * abs(ipo)


@pytest.mark.parametrize("Npts,nit", [(128, 2), (256, 1)])
def test_kdyn_large_grids(Npts, nit):
    """BASELINE configs 3 and 4 at their full grid sizes (few steps: the oracle needs ~1 s per 3-D transform)"""
    from spheremanopt_b200 import kdyn
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0 = kdyn_field(od, 1)
    U = kdyn_field(od, 2)
    Rm, dt = 10.0, 1e-3
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    assert abs(f - fo) <= TOL * abs(fo)
    assert relerr(g[0], go[0]) <= TOL
    assert relerr(g[1], go[1]) <= TOL


@pytest.mark.parametrize("every,adj", [(1, "Discrete"), (4, "Discrete"), (5, "Continuous"), (64, "Discrete")])
def test_kdyn_checkpointed_sweep(every, adj):
    """revolve-style two-level checkpointing (config 4): bit-identical to the fully stored sweep, and equal to the oracle"""
    from spheremanopt_b200 import kdyn
    Npts, nit = 24, 13
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    Rm, dt = 1.0, 1e-3
    full = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, full, "Final", adj)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, full, "Final", adj)
    ck = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=every)
    assert isinstance(ck, kdyn.CheckpointStore) and ck.states_held <= (nit + every - 1) // every + every + 2
    fc = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, ck, "Final", adj)
    gc = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, ck, "Final", adj)
    assert fc == f and np.array_equal(gc[0], g[0]) and np.array_equal(gc[1], g[1])
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Final", adj)
    assert abs(fc - fo) <= TOL * abs(fo) and relerr(gc[0], go[0]) <= TOL and relerr(gc[1], go[1]) <= TOL


def _taylor_slopes(f, grad, ip, X, dX, eps0=1e-3, n=5):
    """second-order Taylor-remainder test (what the reference's Adjoint_Gradient_Test checks, TG:47-150, restated):
    R2(eps) = |f(X + eps dX) - f(X) - eps <Grad_f(X), dX>| must decay like eps^2"""
    f0 = f(X)
    g = grad(X)
    dfd = sum(ip(gi, di) for gi, di in zip(g, dX))
    eps = [eps0 / 2 ** k for k in range(n)]
    R2 = [abs(f([x + e * d for x, d in zip(X, dX)]) - f0 - e * dfd) for e in eps]
    return [np.log(R2[k] / R2[k + 1]) / np.log(2.0) for k in range(n - 1)], R2


def test_taylor_remainder_sh23_config1():
    """BASELINE config 1 at full size (Npts=256, T=50, dt=0.1): slope 2 of the Taylor remainder with the CUDA callables"""
    from spheremanopt_b200 import sh23
    dom, X0 = sh23.Generate_IC(0.0725)
    nit = 500
    store = sh23.GEN_BUFFER(dom, nit)
    args = (dom, 0.1, nit, nit, store)
    dX = np.random.RandomState(5).standard_normal(X0.size) * np.sqrt(0.0725)
    slopes, R2 = _taylor_slopes(lambda X: sh23.FWD_Solve_IVP_Lin(X, *args), lambda X: sh23.ADJ_Solve_IVP_Lin(X, *args),
                                lambda a, b: sh23.Inner_Prod(a, b, dom), [X0], [dX])
    assert all(abs(s - 2.0) < 0.1 for s in slopes), (slopes, R2)


def test_taylor_remainder_kdyn_config3_grid():
    """BASELINE config 3 grid (128^3, Rm=10, dt=1e-3; 20 steps), device-resident vectors, B and U perturbed together"""
    import torch
    from spheremanopt_b200 import kdyn
    from spheremanopt_b200.devvec import DevVec
    Npts, nit = 128, 20
    dom = kdyn.Domain(Npts)
    M = dom.M

    def field(seed):      # band-limited solenoidal field through the library's own transforms
        g = torch.Generator(device="cpu").manual_seed(seed)
        v = torch.randn(3 * M ** 3, dtype=torch.float64, generator=g).to(dom.device)
        c = kdyn.to_coef(dom, v)
        kx, ky, kz = kdyn._wavenumbers(dom)
        k2 = kx * kx + ky * ky + kz * kz
        c = c * torch.exp(-0.2 * torch.sqrt(k2))
        kd = (kx * c[0] + ky * c[1] + kz * c[2]) / torch.where(k2 == 0, torch.ones_like(k2), k2)
        c = torch.stack([c[0] - kx * kd, c[1] - ky * kd, c[2] - kz * kd]) * (k2 != 0)
        v = kdyn.to_grid(dom, c)
        return DevVec(v / np.sqrt(kdyn.Inner_Prod_3(DevVec(v), DevVec(v), dom)))
    X = [field(1), field(2)]
    dX = [field(3), field(4)]
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    args = (dom, 10.0, 1e-3, nit, nit, store)
    slopes, R2 = _taylor_slopes(lambda Y: kdyn.FWD_Solve_IVP_Lin(Y, *args), lambda Y: kdyn.ADJ_Solve_IVP_Lin(Y, *args),
                                lambda a, b: kdyn.Inner_Prod_3(a, b, dom), X, dX, eps0=1e-2)
    assert all(abs(s - 2.0) < 0.1 for s in slopes), (slopes, R2)


def test_kdyn_graph_replay():
    """CUDA-graph replay of the time loops (eager first call, capture on the second, replay afterwards) changes nothing"""
    from spheremanopt_b200 import kdyn
    Npts, nit = 24, 9
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    f0 = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, store)
    g0 = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, store)
    dom.lib.smo_kdyn_use_graph(dom.h, 1)
    for rep in range(4):
        scale = 1.0 + 0.25 * rep     # different inputs through the same graph: nothing but pointers is baked in
        f = kdyn.FWD_Solve_IVP_Lin([scale * B0, U], dom, 1.0, 1e-3, nit, nit, store)
        g = kdyn.ADJ_Solve_IVP_Lin([scale * B0, U], dom, 1.0, 1e-3, nit, nit, store)
        assert abs(f - scale ** 2 * f0) <= 1e-12 * abs(f0) * scale ** 2
        assert relerr(g[0], scale * g0[0]) <= 1e-12 and relerr(g[1], scale ** 2 * g0[1]) <= 1e-12
    ck = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=4)
    for rep in range(3):
        fc = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, ck)
        gc = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, ck)
        assert fc == f0 and np.array_equal(gc[0], g0[0]) and np.array_equal(gc[1], g0[1])


@pytest.mark.parametrize("Npts,nit", [(24, 40), (64, 6)])
def test_kdyn_programmatic_dependent_launch_is_bitwise_neutral(Npts, nit):
    """SMO_OPT_PDL: the kernels of the time loops are launched with programmatic stream serialisation (the next kernel's CTAs
    move in while the previous one drains and block in griddepcontrol.wait); eagerly and from the CUDA graphs, with stored and
    checkpointed sweeps, the results are bit-identical to plain stream order - and equal to the oracle."""
    from spheremanopt_b200 import kdyn, _cabi
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    res = {}
    for pdl in (0, 1):
        assert dom.lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PDL, pdl) == 0
        for graph in (0, 1):
            dom.lib.smo_kdyn_use_graph(dom.h, graph)
            for every in (0, 7):
                store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=every)
                for rep in range(3 if graph else 1):      # eager, capture, replay
                    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, store)
                    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, store)
                    res[(pdl, graph, every, rep)] = (f, g[0].copy(), g[1].copy())
    f0, gb0, gu0 = res[(0, 0, 0, 0)]
    for key, (f, gb, gu) in res.items():
        assert f == f0 and np.array_equal(gb, gb0) and np.array_equal(gu, gu0), key
    if Npts == 24:
        obuf = okd.GEN_BUFFER(Npts, od, nit)
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 1.0, 1e-3, nit, nit, obuf)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 1.0, 1e-3, nit, nit, obuf)
        assert abs(f0 - fo) <= 1e-9 * abs(fo)                     # tolerance of north_star: 1e-9 relative
        assert relerr(gb0, go[0]) <= 1e-9 and relerr(gu0, go[1]) <= 1e-9


@pytest.mark.parametrize("Npts,nit,adj,every", [(24, 12, "Discrete", 0), (32, 6, "Continuous", 0), (24, 12, "Discrete", 5)])
def test_kdyn_integrated_cost(Npts, nit, adj, every):
    """Cost_function="Integrated" (KD:655-669, 738-742, 861-864), stored and checkpointed sweeps"""
    from spheremanopt_b200 import kdyn
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    Rm, dt = 1.0, 1e-3
    store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=every)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store, "Integrated", adj)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store, "Integrated", adj)
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Integrated", adj)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Integrated", adj)
    assert abs(f - fo) <= TOL * abs(fo)
    assert relerr(g[0], go[0]) <= TOL and relerr(g[1], go[1]) <= TOL


def test_kdyn_non_solenoidal_input():
    """adversarial input (not band-limited, not divergence free, non-zero mean): exercises the truncation on first
    gather, the projection of the parameter field U [D2-8] and the k.B carry of the closed-form CNAB1 pencil"""
    from spheremanopt_b200 import kdyn
    Npts, nit = 16, 4
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    r = np.random.RandomState(7)
    B0 = r.standard_normal(od.vec_len * 3)
    U = r.standard_normal(od.vec_len * 3)
    store = kdyn.GEN_BUFFER(Npts, dom, nit)
    f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 2.0, 1e-3, nit, nit, store)
    g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 2.0, 1e-3, nit, nit, store)
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D)
    assert abs(f - fo) <= TOL * abs(fo)
    assert relerr(g[0], go[0]) <= TOL
    assert relerr(g[1], go[1]) <= TOL


def test_kdyn_generate_ic():
    from spheremanopt_b200 import kdyn
    dom, B, U = kdyn.Generate_IC(24, (0., 2. * np.pi), 1.0, True)
    od, Bo, Uo = okd.Generate_IC(24, (0., 2. * np.pi), 1.0, True)
    assert relerr(U, Uo) <= TOL
    assert relerr(B, Bo) <= TOL


def test_devvec_and_sphere_ops():
    import torch
    from spheremanopt_b200 import kdyn
    from spheremanopt_b200.devvec import DevVec
    dom = kdyn.Domain(16)
    od = okd.domain_kdyn(16)
    r = np.random.RandomState(0)
    x, d = r.standard_normal(3 * od.vec_len), r.standard_normal(3 * od.vec_len)
    X, Dv = DevVec.from_numpy(x), DevVec.from_numpy(d)
    ipo = lambda a, b: okd.Inner_Prod_3(a, b, od)
    ip = lambda a, b: kdyn.Inner_Prod_3(a, b, dom)
    assert abs(ip(X, Dv) - ipo(x, d)) <= 1e-12 * abs(ipo(x, x))
    # the optimiser's algebra on opaque device vectors
    y = (np.float64(0.3) * X + Dv * 2.0 - X) .numpy()
    assert relerr(y, 0.3 * x + d * 2.0 - x) <= 1e-15
    assert relerr((-X).numpy(), -x) == 0.0
    c = copy.deepcopy(X)
    assert c.t.data_ptr() != X.t.data_ptr() and relerr(c.numpy(), x) == 0.0
    arr = np.atleast_1d([X, Dv])
    assert arr.dtype == object and arr.shape == (2,)
    ops = dom.vecops(X.n)
    scale = 1.0 / od.M ** 3
    t = ops.project(X.t, Dv.t).cpu().numpy()
    assert relerr(t, osp.tangent_vector(x, d, okd.Inner_Prod_3, (od,))) <= 1e-13
    u = ops.retract(X.t, 0.7, Dv.t, 2.5, scale).cpu().numpy()
    assert relerr(u, osp.Update_vector(x, 0.7, d, 2.5, okd.Inner_Prod_3, (od,))) <= 1e-13


@pytest.mark.parametrize("Npts", [128, 256])
def test_bitwise_reproducible_runs(Npts):
    """the warp-synchronous kernels (one FFT line / field / instance per warp, __syncwarp instead of CTA barriers; TMA copies
    waited on with mbarriers) give bit-identical results run after run - a missing barrier or an early read of a tile still in
    flight would show up as run-to-run noise.  128^3: pair-packed x pass; 256^3: half-length x pass + one-line-per-warp z step.
    (compute-sanitizer's racecheck is not available on this pool.)"""
    import torch
    from spheremanopt_b200 import kdyn, sh23
    from spheremanopt_b200.devvec import DevVec
    dom = kdyn.Domain(Npts)
    g = torch.Generator(device="cpu").manual_seed(11)
    X = [DevVec(torch.randn(3 * dom.M ** 3, dtype=torch.float64, generator=g).to(dom.device)) for _ in range(2)]
    nit = 3 if Npts == 128 else 2
    st = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    runs = []
    for rep in range(3):
        f = kdyn.FWD_Solve_IVP_Lin(X, dom, 10.0, 1e-3, nit, nit, st)
        gr = kdyn.ADJ_Solve_IVP_Lin(X, dom, 10.0, 1e-3, nit, nit, st)
        runs.append((f, gr[0].t.clone(), gr[1].t.clone()))
    for r in runs[1:]:
        assert r[0] == runs[0][0] and torch.equal(r[1], runs[0][1]) and torch.equal(r[2], runs[0][2])
    if Npts != 128:
        return
    sd = sh23.Domain(256)
    for batch in (37, 1300):        # few-instance (latency) variant and ensemble variant (8 CTAs per SM) of the SH23 kernels
        Xs = (torch.randn(batch * sd.M, dtype=torch.float64, generator=g) * 0.05).to(sd.device)
        ss = sh23.GEN_BUFFER(sd, 50, batch=batch)
        ref = None
        for rep in range(3):
            J = sh23.forward_batch(Xs, sd, 0.1, 50, ss).clone()
            G = sh23.adjoint_batch(sd, 0.1, 50, ss).clone()
            if ref is None:
                ref = (J, G)
            assert torch.equal(J, ref[0]) and torch.equal(G, ref[1])


def test_c_abi_host_entry_points():
    """the ctypes stub of INTEGRATION.md, verbatim: numpy in / numpy out through the *_host entry points of the C ABI"""
    import ctypes as C
    from spheremanopt_b200 import _cabi
    lib = _cabi.load()
    od = okd.domain_kdyn(24)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    assert lib.smo_kdyn_create(C.byref(h), 24, 2 * np.pi, 0, 1, None) == 0
    J = C.c_double()
    gB, gU = np.zeros_like(B0), np.zeros_like(U)
    Rm, dt, nit = 1.0, 1e-3, 7
    assert lib.smo_kdyn_forward_host(h, B0.ctypes.data, U.ctypes.data, Rm, dt, nit, None, C.byref(J), 0, None) == 0, lib.smo_last_error()
    assert lib.smo_kdyn_adjoint_host(h, Rm, dt, nit, None, gB.ctypes.data, gU.ctypes.data, 0, None) == 0, lib.smo_last_error()
    D = okd.GEN_BUFFER(24, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    assert abs(-J.value - fo) <= TOL * abs(fo) and relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    assert lib.smo_kdyn_adjoint_host(h, Rm, dt, nit + 1, None, gB.ctypes.data, gU.ctypes.data, 0, None) != 0      # store too small: loud
    lib.smo_kdyn_destroy(h)
    sd = osh.domain_sh23(256)
    X = sh23_input(sd, seed=2)
    hs = C.c_void_p()
    assert lib.smo_sh23_create(C.byref(hs), 256, sd.L, -0.3) == 0
    Js = C.c_double(); G = np.zeros(sd.M)
    assert lib.smo_sh23_forward_host(hs, X.ctypes.data, 1, 0.1, 60, None, C.byref(Js), None) == 0
    assert lib.smo_sh23_adjoint_host(hs, 1, 0.1, 60, None, G.ctypes.data, 0, None) == 0
    Ds = osh.GEN_BUFFER(sd, 60)
    fs = osh.FWD_Solve_IVP_Lin([X], sd, 0.1, 60, 60, Ds)
    assert abs(-Js.value - fs) <= TOL * abs(fs) and relerr(G, osh.ADJ_Solve_IVP_Lin([X], sd, 0.1, 60, 60, Ds)[0]) <= TOL
    lib.smo_sh23_destroy(hs)


def test_grad_f_refills_a_stale_store():
    """Grad_f called for an X other than the one the store was filled for re-runs the forward solve"""
    from spheremanopt_b200 import kdyn, sh23
    od = osh.domain_sh23(64)
    dom = sh23.Domain(64)
    Xa, Xb = sh23_input(od, seed=1), sh23_input(od, seed=2)
    st = sh23.GEN_BUFFER(dom, 20)
    sh23.FWD_Solve_IVP_Lin([Xa], dom, 0.1, 20, 20, st)
    g = sh23.ADJ_Solve_IVP_Lin([Xb], dom, 0.1, 20, 20, st)
    D = osh.GEN_BUFFER(od, 20)
    osh.FWD_Solve_IVP_Lin([Xb], od, 0.1, 20, 20, D)
    assert relerr(g[0], osh.ADJ_Solve_IVP_Lin([Xb], od, 0.1, 20, 20, D)[0]) <= TOL
    okd_ = okd.domain_kdyn(16)
    kd = kdyn.Domain(16)
    Ba, Bb, U = kdyn_field(okd_, 1), kdyn_field(okd_, 3), kdyn_field(okd_, 2)
    ks = kdyn.GEN_BUFFER(16, kd, 4)
    kdyn.FWD_Solve_IVP_Lin([Ba, U], kd, 1.0, 1e-3, 4, 4, ks)
    gk = kdyn.ADJ_Solve_IVP_Lin([Bb, U], kd, 1.0, 1e-3, 4, 4, ks)
    Dk = okd.GEN_BUFFER(16, okd_, 4)
    okd.FWD_Solve_IVP_Lin([Bb, U], okd_, 1.0, 1e-3, 4, 4, Dk)
    go = okd.ADJ_Solve_IVP_Lin([Bb, U], okd_, 1.0, 1e-3, 4, 4, Dk)
    assert relerr(gk[0], go[0]) <= TOL and relerr(gk[1], go[1]) <= TOL


def test_errors_are_loud():
    from spheremanopt_b200 import kdyn, sh23
    with pytest.raises(RuntimeError):
        sh23.Domain(100)          # unsupported size -> error code -> RuntimeError with the library's message
    with pytest.raises(RuntimeError):
        kdyn.Domain(20)
    dom = sh23.Domain(64)
    store = sh23.GEN_BUFFER(dom, 5)
    with pytest.raises(RuntimeError):
        sh23.ADJ_Solve_IVP_Lin([np.zeros(128)], dom, 0.1, 5, 5, store)   # Grad_f before f


def test_kdyn_graph_cache_distinguishes_checkpoint_spacing():
    """two CheckpointStores with the same N_ITERS, the same slot count (N=12: every=4 and every=5 both give 4 slots)
    and - through torch's caching allocator - the same addresses must not replay each other's CUDA graph"""
    import torch
    from spheremanopt_b200 import kdyn
    Npts, nit = 24, 12
    dom = kdyn.Domain(Npts)
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    full = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=0)
    f0 = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, full)
    g0 = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, full)
    del full
    dom.lib.smo_kdyn_use_graph(dom.h, 1)
    ptrs = []
    for every in (4, 5, 4):
        ck = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=every)
        ptrs.append(ck.ptr())
        for rep in range(3):     # eager, capture, replay
            fc = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, ck)
            gc = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, nit, nit, ck)
            assert fc == f0 and np.array_equal(gc[0], g0[0]) and np.array_equal(gc[1], g0[1]), (every, rep)
        del ck
        torch.cuda.synchronize()
    print("checkpoint buffers reused the same address:", len(set(ptrs)) < len(ptrs))


def test_two_domains_on_two_devices_in_one_process():
    """launch configuration (dynamic shared-memory opt-in, occupancy) is per device"""
    import torch
    from spheremanopt_b200 import kdyn
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    od = okd.domain_kdyn(24)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    res = []
    for dev in ("cuda:0", "cuda:1"):
        dom = kdyn.Domain(24, device=dev, distributed=False)
        st = kdyn.GEN_BUFFER(24, dom, 5, checkpoint_every=0)
        f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, 5, 5, st)
        g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, 1.0, 1e-3, 5, 5, st)
        res.append((f, g[0], g[1]))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


def test_fingerprint_sees_single_entry_changes():
    """the identity check of the snapshot store's X is a checksum of EVERY entry of a device vector (round 1 sampled 5)"""
    import torch
    from spheremanopt_b200.devvec import DevVec, fingerprint
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(3 * 36 ** 3, dtype=torch.float64, generator=g).cuda()
    fp = fingerprint(DevVec(x))
    bits = x.cpu().numpy().view(np.uint64)
    with np.errstate(over="ignore"):
        want = int((bits * (2 * np.arange(bits.size, dtype=np.uint64) + np.uint64(1))).sum(dtype=np.uint64))
    assert fp == (x.numel(), want)
    for idx in (1, 777, x.numel() - 2):
        y = x.clone(); y[idx] = y[idx] * (1 + 1e-15) + 1e-300
        assert fingerprint(DevVec(y)) != fp
    perm = x.clone(); perm[[5, 6]] = perm[[6, 5]]
    assert fingerprint(DevVec(perm)) != fp      # position sensitive
