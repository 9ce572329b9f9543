"""world_size-2 CPU test of the slab-decomposed dynamo path: two processes (torch.distributed, gloo) each drive the
host-emulated kernels for their kx / z slab; the all-to-all transposes go through a callback that exchanges the blocks
with gloo.  Checks J, Grad_f and the transforms against the single-process oracle.  Test infrastructure only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, Npts, nit, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from oracle import kdyn as okd
    from tests.common import kdyn_field
    from tests.emul import emul
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = emul.lib()
        A2A = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)

        def a2a(send, recv, nbytes, user):
            src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_ubyte)), shape=(world * nbytes,))
            dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_ubyte)), shape=(world * nbytes,))
            mine = torch.from_numpy(src.copy())
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            for s in range(world):   # block `rank` of what peer s sent
                dst[s * nbytes:(s + 1) * nbytes] = parts[s].numpy()[rank * nbytes:(rank + 1) * nbytes]
        cb = A2A(a2a)
        L.smo_emul_make_comm.restype = C.c_void_p
        L.smo_emul_make_comm.argtypes = [A2A, C.c_void_p]
        comm = L.smo_emul_make_comm(cb, None)
        od = okd.domain_kdyn(Npts)
        M = od.M
        nz = M // world
        z0 = rank * nz
        B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
        slab = lambda v: np.ascontiguousarray(v.reshape(3, M, M, M)[:, :, :, z0:z0 + nz]).reshape(-1)
        h = C.c_void_p()
        emul.check(L.smo_kdyn_create(C.byref(h), Npts, od.L, rank, world, comm))
        gsz = L.smo_kdyn_grid_elems(h)
        assert gsz == M * M * nz
        Bs, Us = slab(B0), slab(U)
        # transforms: round trip through the distributed coefficient layout
        coef = np.zeros(3 * L.smo_kdyn_coef_elems(h), dtype=complex)
        emul.check(L.smo_kdyn_to_coef(h, emul.ptr(Bs), emul.ptr(coef), None))
        nkx = od.Nh // world
        cc = coef.reshape(3, nkx, od.Nc, od.Nc + 1)[..., :od.Nc]
        want = np.stack([od.to_coef_3d(x) for x in okd.Vec_to_Field(od, B0)])[:, rank * nkx:(rank + 1) * nkx]
        e_coef = float(np.abs(cc - want).max() / np.abs(want).max())
        snaps = np.zeros(L.smo_kdyn_snapshot_bytes(h, nit) // 16, dtype=complex)
        J = C.c_double()
        emul.check(L.smo_kdyn_forward(h, emul.ptr(Bs), emul.ptr(Us), 1.5, 1e-3, nit, emul.ptr(snaps), C.byref(J), 0, None))
        gB, gU = np.zeros(3 * gsz), np.zeros(3 * gsz)
        emul.check(L.smo_kdyn_adjoint(h, 1.5, 1e-3, nit, emul.ptr(snaps), emul.ptr(gB), emul.ptr(gU), 0, None))
        Jt = torch.tensor([J.value], dtype=torch.float64)
        dist.all_reduce(Jt)
        D = okd.GEN_BUFFER(Npts, od, nit)
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, 1.5, 1e-3, nit, nit, D)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, 1.5, 1e-3, nit, nit, D)
        eJ = abs(-Jt.item() - fo) / abs(fo)
        eB = float(np.abs(gB - slab(go[0])).max() / np.abs(go[0]).max())
        eU = float(np.abs(gU - slab(go[1])).max() / np.abs(go[1]).max())
        # cost "Integrated" through the checkpointed sweep (every = 2): per-rank partial sums of |B^n|^2 and segment recomputation
        ck = np.zeros(L.smo_kdyn_checkpoint_bytes(h, nit, 2) // 16, dtype=complex)
        seg = np.zeros(L.smo_kdyn_segment_bytes(h, 2) // 16, dtype=complex)
        emul.check(L.smo_kdyn_forward_ckpt(h, emul.ptr(Bs), emul.ptr(Us), 1.5, 1e-3, nit, 2, emul.ptr(ck), C.byref(J), 2, None))
        emul.check(L.smo_kdyn_adjoint_ckpt(h, 1.5, 1e-3, nit, 2, emul.ptr(ck), emul.ptr(seg), emul.ptr(gB), emul.ptr(gU), 2, None))
        Jt = torch.tensor([J.value], dtype=torch.float64)
        dist.all_reduce(Jt)
        fi = okd.FWD_Solve_IVP_Lin([B0, U], od, 1.5, 1e-3, nit, nit, D, "Integrated")
        gi = okd.ADJ_Solve_IVP_Lin([B0, U], od, 1.5, 1e-3, nit, nit, D, "Integrated")
        eJ = max(eJ, abs(-Jt.item() - fi) / abs(fi))
        eB = max(eB, float(np.abs(gB - slab(gi[0])).max() / np.abs(gi[0]).max()))
        eU = max(eU, float(np.abs(gU - slab(gi[1])).max() / np.abs(gi[1]).max()))
        q.put((rank, e_coef, eJ, eB, eU))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Npts,nit", [(16, 3), (24, 2)])
def test_two_rank_slab_decomposition(Npts, nit):
    import torch.multiprocessing as mp
    from tests.emul import emul
    emul.lib()   # build once, before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + Npts
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Npts, nit, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    for _ in procs:
        res.append(q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_coef, eJ, eB, eU in res:
        assert e_coef < 1e-12 and eJ < 1e-11 and eB < 1e-11 and eU < 1e-11, (rank, e_coef, eJ, eB, eU)
