"""CPU check of the kernels' index logic: the SAME phase bodies that nvcc compiles for sm_100a are compiled by g++
(-DSMO_EMUL) and run thread-by-thread on the host, then compared with the oracle.  The emulation library is test
infrastructure (tests/emul) and is never loaded by the product package."""
import ctypes as C

import numpy as np
import pytest

from oracle import kdyn as okd
from oracle import sh23 as osh
from oracle import sphere as osp
from tests.common import kdyn_field, relerr, sh23_input
from tests.emul import emul

TOL = 1e-11


@pytest.fixture(scope="module")
def L():
    return emul.lib()


@pytest.mark.parametrize("Npts", [64, 128, 256])
def test_sh23_emulated(L, Npts):
    od = osh.domain_sh23(Npts)
    batch, dt, nit = 6, 0.1, 12        # 6 instances: 4 per CTA -> one full and one ragged CTA
    X = np.stack([sh23_input(od, seed=b, amp=0.04 + 0.003 * b) for b in range(batch)])
    h = C.c_void_p()
    emul.check(L.smo_sh23_create(C.byref(h), Npts, od.L, -0.3))
    nb = L.smo_sh23_snapshot_bytes(h, nit) // 8      # opaque store: [nit+1][M] grid values + Nh final coefficients per instance
    snaps = np.zeros((batch, nb))
    J = np.zeros(batch); G = np.zeros((batch, od.M)); Gc = np.zeros((batch, od.M))
    emul.check(L.smo_sh23_forward(h, emul.ptr(X), batch, dt, nit, emul.ptr(snaps), emul.ptr(J), None))
    emul.check(L.smo_sh23_adjoint(h, batch, dt, nit, emul.ptr(snaps), emul.ptr(G), 0, None))
    emul.check(L.smo_sh23_adjoint(h, batch, dt, nit, emul.ptr(snaps), emul.ptr(Gc), 1, None))
    P = np.zeros((batch, od.M))
    emul.check(L.smo_sh23_prep(h, emul.ptr(X), batch, 0.01, 5, emul.ptr(P), None))
    A = np.zeros((nit + 1, batch, od.Nh), dtype=complex)
    for n in range(nit + 1):
        emul.check(L.smo_sh23_snapshot_coef(h, emul.ptr(snaps), batch, nit, n, emul.ptr(A[n]), None))
    for b in range(batch):
        D = osh.GEN_BUFFER(od, nit)
        fo = osh.FWD_Solve_IVP_Lin([X[b]], od, dt, nit, nit, D)
        assert abs(-J[b] - fo) <= TOL * abs(fo)
        assert relerr(A[:, b, :].T, D['A_fwd']) <= TOL
        assert relerr(snaps[b, 3 * od.M:4 * od.M], od.to_grid_1d(D['A_fwd'][:, 3])) <= TOL      # the store holds grid values
        assert relerr(G[b], osh.ADJ_Solve_IVP_Lin([X[b]], od, dt, nit, nit, D)[0]) <= TOL
        assert relerr(Gc[b], osh.ADJ_Solve_IVP_Lin([X[b]], od, dt, nit, nit, D, None, "Continuous")[0]) <= TOL
        assert relerr(P[b], osh.FWD_Solve_IVP_PREP(X[b], od, 0.01, 5)) <= TOL
    L.smo_sh23_destroy(h)


@pytest.mark.parametrize("Npts,nit", [(16, 4), (24, 3), (32, 2)])
def test_kdyn_emulated(L, Npts, nit):
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    if Npts == 16:   # adversarial input: not band limited, not solenoidal, non-zero mean
        r = np.random.RandomState(3)
        B0, U = r.standard_normal(B0.size), r.standard_normal(U.size)
    h = C.c_void_p()
    emul.check(L.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    csz, gsz = L.smo_kdyn_coef_elems(h), L.smo_kdyn_grid_elems(h)
    assert gsz == od.M ** 3
    coef = np.zeros((3, csz), dtype=complex)
    emul.check(L.smo_kdyn_to_coef(h, emul.ptr(B0), emul.ptr(coef), None))
    cc = coef.reshape(3, od.Nh, od.Nc, od.Nc + 1)[..., :od.Nc]
    want = np.stack([od.to_coef_3d(x) for x in okd.Vec_to_Field(od, B0)])
    assert relerr(cc, want) <= TOL
    if Npts == 32:   # z-chunked y -> x -> y sequence (the L2-residency schedule used at 128^3)
        emul.check(L.smo_kdyn_set_chunks(h, 3, 2))
    snaps = np.zeros(L.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
    J = C.c_double()
    Rm, dt = 2.0, 1e-3
    emul.check(L.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), Rm, dt, nit, emul.ptr(snaps), C.byref(J), 0, None))
    gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
    emul.check(L.smo_kdyn_adjoint(h, Rm, dt, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), 0, None))
    D = okd.GEN_BUFFER(Npts, od, nit)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
    assert abs(-J.value - fo) <= TOL * abs(fo)
    sc = np.zeros((3, csz), dtype=complex)
    for n in (0, 1, nit):      # stored states (x-spectral form) converted back to coefficients
        emul.check(L.smo_kdyn_snapshot_coef(h, emul.ptr(snaps), nit, n, emul.ptr(sc), None))
        for c, key in enumerate(('A_fwd', 'B_fwd', 'C_fwd')):
            assert relerr(sc[c].reshape(od.Nh, od.Nc, od.Nc + 1)[..., :od.Nc], D[key][..., n]) <= TOL
    assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    if Npts == 24:
        gBc, gUc = np.zeros(3 * gsz), np.zeros(3 * gsz)
        emul.check(L.smo_kdyn_adjoint(h, Rm, dt, nit, emul.ptr(snaps), emul.ptr(gBc), emul.ptr(gUc), 1, None))
        goc = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Final", "Continuous")
        assert relerr(gBc, goc[0]) <= TOL and relerr(gUc, goc[1]) <= TOL
        out = np.zeros(3 * gsz)
        emul.check(L.smo_kdyn_prep(h, emul.ptr(B0), emul.ptr(U), Rm, dt, 2, emul.ptr(out), None))
        Bc = okd.FWD_Solve_IVP_Prep(B0, U, od, Rm, dt, 2)
        assert relerr(out, okd.Field_to_Vec(od, *[od.to_grid_3d(c) for c in Bc])) <= TOL
    L.smo_kdyn_destroy(h)


@pytest.mark.parametrize("every,cont", [(1, 0), (3, 0), (4, 1), (7, 0), (16, 1)])
def test_kdyn_checkpointed_emulated(L, every, cont):
    """two-level checkpointing must reproduce the fully stored sweep bit for bit (same kernels, same order)"""
    Npts, nit = 16, 7
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    emul.check(L.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    gsz = L.smo_kdyn_grid_elems(h)
    Rm, dt = 2.0, 1e-3
    snaps = np.zeros(L.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
    J = C.c_double(); Jc = C.c_double()
    emul.check(L.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), Rm, dt, nit, emul.ptr(snaps), C.byref(J), 0, None))
    gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
    emul.check(L.smo_kdyn_adjoint(h, Rm, dt, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), cont, None))
    ck = np.zeros(L.smo_kdyn_checkpoint_bytes(h, nit, every) // 16, dtype=complex)
    seg = np.zeros(L.smo_kdyn_segment_bytes(h, every) // 16, dtype=complex)
    emul.check(L.smo_kdyn_forward_ckpt(h, emul.ptr(B0), emul.ptr(U), Rm, dt, nit, every, emul.ptr(ck), C.byref(Jc), 0, None))
    gBc, gUc = np.zeros(3 * gsz), np.zeros(3 * gsz)
    emul.check(L.smo_kdyn_adjoint_ckpt(h, Rm, dt, nit, every, emul.ptr(ck), emul.ptr(seg), emul.ptr(gBc), emul.ptr(gUc), cont, None))
    assert Jc.value == J.value
    assert np.array_equal(gBc, gB) and np.array_equal(gUc, gU)
    L.smo_kdyn_destroy(h)


@pytest.mark.parametrize("adj,every", [(0, 0), (1, 0), (0, 3)])
def test_kdyn_integrated_cost_emulated(L, adj, every):
    """Cost_function="Integrated" (KD:655-669, 738-742, 861-864): J = dt sum_n <B^n,B^n>, adjoint source -2 B_f"""
    Npts, nit = 16, 5
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    emul.check(L.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    gsz = L.smo_kdyn_grid_elems(h)
    Rm, dt = 2.0, 1e-3
    J = C.c_double()
    gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
    flags = 2 | adj      # SMO_COST_INTEGRATED | SMO_ADJOINT_CONTINUOUS
    if every:
        ck = np.zeros(L.smo_kdyn_checkpoint_bytes(h, nit, every) // 16, dtype=complex)
        seg = np.zeros(L.smo_kdyn_segment_bytes(h, every) // 16, dtype=complex)
        emul.check(L.smo_kdyn_forward_ckpt(h, emul.ptr(B0), emul.ptr(U), Rm, dt, nit, every, emul.ptr(ck), C.byref(J), flags, None))
        emul.check(L.smo_kdyn_adjoint_ckpt(h, Rm, dt, nit, every, emul.ptr(ck), emul.ptr(seg), emul.ptr(gB), emul.ptr(gU), flags, None))
    else:
        snaps = np.zeros(L.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
        emul.check(L.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), Rm, dt, nit, emul.ptr(snaps), C.byref(J), flags, None))
        emul.check(L.smo_kdyn_adjoint(h, Rm, dt, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), flags, None))
    D = okd.GEN_BUFFER(Npts, od, nit)
    at = "Continuous" if adj else "Discrete"
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Integrated", at)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D, "Integrated", at)
    assert abs(-J.value - fo) <= TOL * abs(fo)
    assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    L.smo_kdyn_destroy(h)


def test_host_buffer_entry_points_emulated(L):
    """the *_host forms of the C ABI (what INTEGRATION.md's ctypes stub binds): handle-owned device buffers and snapshot store"""
    od = okd.domain_kdyn(16)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    emul.check(L.smo_kdyn_create(C.byref(h), 16, od.L, 0, 1, None))
    J = C.c_double()
    gB, gU = np.zeros_like(B0), np.zeros_like(U)
    emul.check(L.smo_kdyn_forward_host(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, 3, None, C.byref(J), 0, None))
    emul.check(L.smo_kdyn_adjoint_host(h, 2.0, 1e-3, 3, None, emul.ptr(gB), emul.ptr(gU), 0, None))
    D = okd.GEN_BUFFER(16, od, 3)
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, 3, 3, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, 3, 3, D)
    assert abs(-J.value - fo) <= TOL * abs(fo) and relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    out = np.zeros_like(B0)
    emul.check(L.smo_kdyn_prep_host(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, 2, emul.ptr(out), None))
    Bc = okd.FWD_Solve_IVP_Prep(B0, U, od, 2.0, 1e-3, 2)
    assert relerr(out, okd.Field_to_Vec(od, *[od.to_grid_3d(c) for c in Bc])) <= TOL
    L.smo_kdyn_destroy(h)
    sd = osh.domain_sh23(64)
    X = sh23_input(sd, seed=2)
    hs = C.c_void_p()
    emul.check(L.smo_sh23_create(C.byref(hs), 64, sd.L, -0.3))
    Js = C.c_double(); G = np.zeros(sd.M)
    emul.check(L.smo_sh23_forward_host(hs, emul.ptr(X), 1, 0.1, 10, None, C.byref(Js), None))
    emul.check(L.smo_sh23_adjoint_host(hs, 1, 0.1, 10, None, emul.ptr(G), 0, None))
    Ds = osh.GEN_BUFFER(sd, 10)
    fs = osh.FWD_Solve_IVP_Lin([X], sd, 0.1, 10, 10, Ds)
    assert abs(-Js.value - fs) <= TOL * abs(fs) and relerr(G, osh.ADJ_Solve_IVP_Lin([X], sd, 0.1, 10, 10, Ds)[0]) <= TOL
    L.smo_sh23_destroy(hs)


def test_kdyn_radix24_paths_emulated():
    """the code paths of the 256^3 grid (M = 384 = 24 x 16: 24 stage threads per FFT, one FFT per warp in the fused x pass,
    CTA-wide barriers in the y / z kernels) at a size the host emulation can afford: M = 48 factorised as 24 x 2"""
    L24 = emul.lib_variant("r24", ["SMO_TEST_R24"])
    Npts, nit = 32, 3
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    _cabi_check = lambda rc: emul._cabi.check(L24, rc)
    _cabi_check(L24.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    gsz = L24.smo_kdyn_grid_elems(h)
    D = okd.GEN_BUFFER(Npts, od, nit)
    for cost, flag in (("Final", 0), ("Integrated", 2)):
        snaps = np.zeros(L24.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
        J = C.c_double()
        gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
        _cabi_check(L24.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, nit, emul.ptr(snaps), C.byref(J), flag, None))
        _cabi_check(L24.smo_kdyn_adjoint(h, 2.0, 1e-3, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), flag, None))
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        assert abs(-J.value - fo) <= TOL * abs(fo)
        assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    L24.smo_kdyn_destroy(h)


def test_vector_kernels_emulated(L):
    od = okd.domain_kdyn(16)
    n = 3 * od.M ** 3
    r = np.random.RandomState(0)
    x, d = r.standard_normal(n), r.standard_normal(n)
    work = np.zeros(L.smo_vec_work_bytes(n) // 8 + 1)
    scale = 1.0 / od.M ** 3
    out = C.c_double()
    emul.check(L.smo_vec_dot(emul.ptr(x), emul.ptr(d), n, scale, C.byref(out), emul.ptr(work), None))
    assert abs(out.value - okd.Inner_Prod_3(x, d, od)) <= 1e-13 * okd.Inner_Prod_3(x, x, od)
    y = np.zeros(n)
    emul.check(L.smo_vec_axpby(0.3, emul.ptr(x), -2.0, emul.ptr(d), emul.ptr(y), n, None))
    assert relerr(y, 0.3 * x - 2.0 * d) <= 1e-15
    emul.check(L.smo_vec_project(emul.ptr(x), emul.ptr(d), emul.ptr(y), n, emul.ptr(work), None))
    assert relerr(y, osp.tangent_vector(x, d, okd.Inner_Prod_3, (od,))) <= 1e-13
    emul.check(L.smo_vec_retract(emul.ptr(x), 0.7, emul.ptr(d), 2.5, scale, emul.ptr(y), n, emul.ptr(work), None))
    assert relerr(y, osp.Update_vector(x, 0.7, d, 2.5, okd.Inner_Prod_3, (od,))) <= 1e-13
    # ragged length (not a multiple of the chunk) and tiny vectors
    for m in (1, 511, 8193):
        emul.check(L.smo_vec_dot(emul.ptr(x), emul.ptr(d), m, 1.0, C.byref(out), emul.ptr(work), None))
        assert abs(out.value - float(np.dot(x[:m], d[:m]))) <= 1e-12 * m


def test_vector_kernels_alignment_paths_and_checksum_emulated(L):
    """128-bit (pair) path vs the scalar path (operands offset by one double), odd lengths; the 64-bit checksum"""
    r = np.random.RandomState(1)
    for n in (2, 3, 8192, 8193, 20001):
        buf = [r.standard_normal(n + 2) for _ in range(3)]
        work = np.zeros(L.smo_vec_work_bytes(n) // 8 + 1)
        for off in (0, 1):       # off = 1: 8-byte aligned only -> scalar path
            x, d, y = (b[off:off + n] for b in buf)
            out = C.c_double()
            emul.check(L.smo_vec_dot(emul.ptr(x), emul.ptr(d), n, 1.0, C.byref(out), emul.ptr(work), None))
            assert abs(out.value - float(np.dot(x, d))) <= 1e-12 * n
            emul.check(L.smo_vec_axpby(0.3, emul.ptr(x), -2.0, emul.ptr(d), emul.ptr(y), n, None))
            assert relerr(y, 0.3 * x - 2.0 * d) <= 1e-15
            emul.check(L.smo_vec_project(emul.ptr(x), emul.ptr(d), emul.ptr(y), n, emul.ptr(work), None))
            assert relerr(y, d - (np.dot(x, d) / np.dot(x, x)) * x) <= 1e-12
            emul.check(L.smo_vec_retract(emul.ptr(x), 0.7, emul.ptr(d), 2.5, 0.5, emul.ptr(y), n, emul.ptr(work), None))
            f = x + 0.7 * d
            assert relerr(y, f * np.sqrt(2.5 / (0.5 * np.dot(f, f)))) <= 1e-12
        x = np.ascontiguousarray(buf[0][:n])
        h = C.c_ulonglong()
        emul.check(L.smo_vec_checksum(emul.ptr(x), n, C.byref(h), emul.ptr(work), None))
        bits = x.view(np.uint64)
        with np.errstate(over="ignore"):
            assert h.value == int((bits * (2 * np.arange(n, dtype=np.uint64) + np.uint64(1))).sum(dtype=np.uint64))


def test_kdyn_half_length_x_pass_emulated():
    """the half-length fused x pass of the 256^3 grid (xpass_half.cuh: one real column = one complex FFT of length M/2) at a
    size the host emulation can afford: M = 96 served by XFusedH<Fac<8,6>> (test-only switch SMO_TEST_HALFX), both costs"""
    LH = emul.lib_variant("halfx", ["SMO_TEST_HALFX"])
    Npts, nit = 64, 2
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    chk = lambda rc: emul._cabi.check(LH, rc)
    chk(LH.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    gsz = LH.smo_kdyn_grid_elems(h)
    D = okd.GEN_BUFFER(Npts, od, nit)
    for cost, flag in (("Final", 0), ("Integrated", 2)):
        snaps = np.zeros(LH.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
        J = C.c_double()
        gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
        chk(LH.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, nit, emul.ptr(snaps), C.byref(J), flag, None))
        chk(LH.smo_kdyn_adjoint(h, 2.0, 1e-3, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), flag, None))
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        assert abs(-J.value - fo) <= TOL * abs(fo)
        assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    LH.smo_kdyn_destroy(h)


@pytest.mark.parametrize("variant,Npts,nit,every", [("plain", 16, 5, 0), ("plain", 24, 7, 3), ("halfx", 64, 2, 0)])
def test_kdyn_grid_accumulation_emulated(variant, Npts, nit, every):
    """SMO_OPT_GRID_ACC: the adjoint x pass sums (curl G) x B_f on the real grid (tile-major) and the r2c transform runs once
    after the sweep - stored and checkpointed sweeps, both x-pass variants, both costs"""
    Lv = emul.lib() if variant == "plain" else emul.lib_variant("halfx", ["SMO_TEST_HALFX"])
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    chk = lambda rc: emul._cabi.check(Lv, rc)
    chk(Lv.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    chk(Lv.smo_kdyn_set_option(h, emul._cabi.SMO_OPT_GRID_ACC, 1))
    gsz = Lv.smo_kdyn_grid_elems(h)
    D = okd.GEN_BUFFER(Npts, od, nit)
    for cost, flag in (("Final", 0), ("Integrated", 2)):
        J = C.c_double()
        gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
        if every:
            ck = np.zeros(Lv.smo_kdyn_checkpoint_bytes(h, nit, every) // 16, dtype=complex)
            seg = np.zeros(Lv.smo_kdyn_segment_bytes(h, every) // 16, dtype=complex)
            chk(Lv.smo_kdyn_forward_ckpt(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, nit, every, emul.ptr(ck), C.byref(J), flag, None))
            chk(Lv.smo_kdyn_adjoint_ckpt(h, 2.0, 1e-3, nit, every, emul.ptr(ck), emul.ptr(seg), emul.ptr(gB), emul.ptr(gU), flag, None))
        else:
            snaps = np.zeros(Lv.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
            chk(Lv.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, nit, emul.ptr(snaps), C.byref(J), flag, None))
            chk(Lv.smo_kdyn_adjoint(h, 2.0, 1e-3, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), flag, None))
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D, cost)
        assert abs(-J.value - fo) <= TOL * abs(fo)
        assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    Lv.smo_kdyn_destroy(h)


@pytest.mark.parametrize("variant,Npts,nit", [("plain", 24, 4), ("halfx", 64, 2)])
def test_kdyn_bulk_velocity_tile_emulated(variant, Npts, nit):
    """SMO_OPT_BULK_U: the velocity tile lives in HBM in its swizzled shared-memory order and arrives as one contiguous (TMA bulk)
    copy - checks the pre-swizzled layout of UTile / UTileH against the ui() the x passes read with (together with grid accumulation)"""
    Lv = emul.lib() if variant == "plain" else emul.lib_variant("halfx", ["SMO_TEST_HALFX"])
    od = okd.domain_kdyn(Npts)
    B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
    h = C.c_void_p()
    chk = lambda rc: emul._cabi.check(Lv, rc)
    chk(Lv.smo_kdyn_create(C.byref(h), Npts, od.L, 0, 1, None))
    chk(Lv.smo_kdyn_set_option(h, emul._cabi.SMO_OPT_BULK_U, 1))
    chk(Lv.smo_kdyn_set_option(h, emul._cabi.SMO_OPT_GRID_ACC, 1))
    gsz = Lv.smo_kdyn_grid_elems(h)
    D = okd.GEN_BUFFER(Npts, od, nit)
    snaps = np.zeros(Lv.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
    J = C.c_double()
    gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
    chk(Lv.smo_kdyn_forward(h, emul.ptr(B0), emul.ptr(U), 2.0, 1e-3, nit, emul.ptr(snaps), C.byref(J), 0, None))
    chk(Lv.smo_kdyn_adjoint(h, 2.0, 1e-3, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), 0, None))
    fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D)
    go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 2.0, 1e-3, nit, nit, D)
    assert abs(-J.value - fo) <= TOL * abs(fo)
    assert relerr(gB, go[0]) <= TOL and relerr(gU, go[1]) <= TOL
    Lv.smo_kdyn_destroy(h)
