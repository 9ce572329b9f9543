"""B200 drop-in for Example_Problems/Periodic_Domain(Fourier)/Kinematic_Dynamo/FWD_Solve_KDyn.py (alias ``KD``).

Same function names, argument order and return conventions as the reference, so that its driver block
(KD:1025-1067) works unchanged with these callables:

    domain, Bx0, Ux = Generate_IC(Npts, X_domain, M_0, Noise)         # KD:183-317
    X_FWD_DICT      = GEN_BUFFER(Npts, domain, N_SUB_ITERS)            # KD:319-355
    args_IP = (domain, None)
    args_f  = [domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, "Final", "Discrete"]
    Optimise_On_Multi_Sphere([Bx0, Ux], [M_0, E_0], FWD_Solve_IVP_Lin, ADJ_Solve_IVP_Lin, Inner_Prod_3, args_f, args_IP, ...)

Vectors are float64 numpy arrays ``concat(Fx.ravel(), Fy.ravel(), Fz.ravel())`` on the 3/2-dealiased M^3 grid
(KD:137; "Mode H", replicated on every rank exactly like the reference's allgather'ed vectors) or ``DevVec`` objects
holding this rank's z-slab [3][M][M][nz] in HBM ("Mode D").  With more than one rank (one process per GPU,
torch.distributed initialised) the grid is slab-decomposed: coefficient space along kx, grid space along z, with
all-to-all transposes inside the library - the replacement of Dedalus' MPI transposes [D2-9].
All arithmetic happens in libsmo_b200.so; there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi, sideout
from .devvec import DevVec, VecOps, _stream_ptr, fingerprint

SIDE_OUTPUTS = False   # True (or SMO_SIDE_OUTPUTS=1): write the reference's CheckPoints / scalar_data handlers (sideout.py)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


class Domain:
    """Stand-in for the dedalus domain of KD:212-216 (three Fourier bases, Npts modes each, dealias 3/2)."""

    def __init__(self, Npts, X=(0., 2. * np.pi), device=None, distributed=None, peer_memory=True):
        self.lib = _cabi.load()
        self.N = int(Npts)
        self.dealias = 3 / 2
        self.M = 3 * self.N // 2
        self.Nh = self.N // 2
        self.kmax = (self.N - 1) // 2
        self.Nc = 2 * self.kmax + 1
        self.interval = (float(X[0]), float(X[1]))
        self.L = self.interval[1] - self.interval[0]
        self.hypervolume = self.L ** 3
        dist = _dist() if distributed in (None, True) else None
        self.rank = dist.get_rank() if dist else 0
        self.nranks = dist.get_world_size() if dist else 1
        if device is None:
            device = "cuda:%d" % (self.rank % max(torch.cuda.device_count(), 1))
        self.device = torch.device(device)
        self.comm = None
        with torch.cuda.device(self.device):
            if self.nranks > 1:
                nb = self.lib.smo_comm_unique_id_bytes()
                idbuf = (C.c_ubyte * nb)()
                if self.rank == 0:
                    _cabi.check(self.lib, self.lib.smo_comm_get_unique_id(idbuf))
                t = torch.tensor(list(idbuf), dtype=torch.uint8, device=self.device)
                dist.broadcast(t, 0)
                idbuf = (C.c_ubyte * nb)(*t.cpu().tolist())
                comm = C.c_void_p()
                _cabi.check(self.lib, self.lib.smo_comm_create(C.byref(comm), idbuf, self.nranks, self.rank))
                self.comm = comm
            h = C.c_void_p()
            _cabi.check(self.lib, self.lib.smo_kdyn_create(C.byref(h), self.N, self.L, self.rank, self.nranks, self.comm))
        self.h = h
        # development / A-B switches: SMO_KDYN_OPTS="key=value,..." applies smo_kdyn_set_option to every new Domain
        for kv in filter(None, os.environ.get("SMO_KDYN_OPTS", "").split(",")):
            k, v = kv.split("=")
            _cabi.check(self.lib, self.lib.smo_kdyn_set_option(h, int(k), int(v)))
        self.peer = False
        if self.nranks > 1 and peer_memory:
            # fused transposes: exchange CUDA IPC handles of the pencil buffers so that the FFT passes store straight
            # into the peers' memory over NVLink (falls back to grouped ncclSend/ncclRecv if IPC is unavailable)
            with torch.cuda.device(self.device):
                nb = self.lib.smo_kdyn_peer_handle_bytes()
                mine = (C.c_ubyte * nb)()
                ok = self.lib.smo_kdyn_peer_export(h, mine) == 0
                t = torch.tensor(list(mine) + [1 if ok else 0], dtype=torch.uint8, device=self.device)
                parts = [torch.empty_like(t) for _ in range(self.nranks)]
                dist.all_gather(parts, t)
                allb = torch.stack(parts).cpu()
                if bool(allb[:, -1].all()):
                    blob = (C.c_ubyte * (nb * self.nranks))(*allb[:, :-1].reshape(-1).tolist())
                    ok = self.lib.smo_kdyn_peer_attach(h, blob) == 0
                else:
                    ok = False
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) != 1:
                    raise RuntimeError("peer-memory attachment failed on some rank: %s" % self.lib.smo_last_error().decode())
                self.peer = True
        self.nz = self.M // self.nranks
        self.z0 = self.rank * self.nz
        self.nkx = self.Nh // self.nranks
        self.gsize = self.lib.smo_kdyn_grid_elems(h)     # M*M*nz
        self.csize = self.lib.smo_kdyn_coef_elems(h)     # nkx*Nc*(Nc+1)
        self.vec_len = 3 * self.M ** 3                   # the reference's vector length
        self._vecops = {}

    def vecops(self, n):
        if n not in self._vecops:
            self._vecops[n] = VecOps(n, self.device)
        return self._vecops[n]

    # host vector <-> local slab -----------------------------------------------------------------------------
    def slab_from_host(self, x):
        """full reference vector (3*M^3) -> this rank's device slab [3][M][M][nz]"""
        M = self.M
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if a.size != 3 * M ** 3:
            raise ValueError("vector of %d entries, expected 3*M^3 = %d" % (a.size, 3 * M ** 3))
        if self.nranks == 1:
            return torch.from_numpy(a).to(self.device, non_blocking=True)
        # several ranks: one strided 2-D copy straight out of the caller's array (Vec_to_Field's local slicing, KD:156-169)
        t = torch.empty(3 * self.gsize, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib, self.lib.smo_kdyn_slab_copy(self.h, t.data_ptr(), a.ctypes.data, None, _stream_ptr()))
            torch.cuda.current_stream().synchronize()     # (`a` may be a temporary)
        return t

    def host_from_slab(self, t):
        """device slab -> full reference vector on the host (all-gather over ranks, like KD:118-137)"""
        M = self.M
        if self.nranks > 1:
            dist = _dist()
            parts = [torch.empty_like(t) for _ in range(self.nranks)]
            dist.all_gather(parts, t)
            t = torch.cat([p.view(3, M, M, self.nz) for p in parts], dim=3).reshape(-1)
        # D2H through page-locked memory (torch's caching host allocator recycles the blocks): ~50 GB/s instead of the
        # pageable path's ~6 GB/s; the numpy array returned to the optimiser keeps the pinned block alive
        out = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        out.copy_(t)
        return out.numpy()

    def allreduce_sum(self, v):
        if self.nranks == 1:
            return v
        dist = _dist()
        # gather the per-rank partials and add them in rank order: bit-identical on every rank
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        parts = [torch.empty_like(t) for _ in range(self.nranks)]
        dist.all_gather(parts, t)
        s = 0.0
        for p in torch.cat(parts).cpu().tolist():
            s += p
        return s

    def all_agree(self, flag, op):
        """a yes/no decision every rank must take identically: MIN ("all say yes") or MAX ("any says yes") over the ranks"""
        if self.nranks == 1:
            return bool(flag)
        dist = _dist()
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.MAX)
        return bool(int(t.item()))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.smo_kdyn_destroy(self.h)
                self.h = None
            if getattr(self, "comm", None):
                self.lib.smo_comm_destroy(self.comm)
                self.comm = None
        except Exception:
            pass


class SnapshotStore(dict):
    """Device-resident replacement of ``{'A_fwd','B_fwd','C_fwd'}`` (KD:347-355).  Opaque HBM store: every forward state is
    kept as x-spectra on this rank's z-slab ([N_SUB_ITERS+1][3][Npts/2][M][nz] complex128: what the adjoint x pass reads,
    written for free by the forward solve's y pass) plus the coefficients of the final state.  ``['A_fwd']`` etc. convert
    back and return host copies in the reference's [Npts/2, Npts-1, Npts-1, N_SUB_ITERS+1] orientation (single rank
    only), for inspection."""

    def __init__(self, domain, n_iters):
        super().__init__()
        self.domain, self.n_iters = domain, int(n_iters)
        nbytes = domain.lib.smo_kdyn_snapshot_bytes(domain.h, self.n_iters)
        self.buf = torch.zeros(nbytes // 16, dtype=torch.complex128, device=domain.device)
        self.valid = False

    def ptr(self):
        return self.buf.data_ptr()

    def __getitem__(self, key):
        c = {'A_fwd': 0, 'B_fwd': 1, 'C_fwd': 2}[key]
        d = self.domain
        out = np.zeros((d.nkx, d.Nc, d.Nc, self.n_iters + 1), dtype=np.complex128)
        coef = torch.zeros(3 * d.csize, dtype=torch.complex128, device=d.device)
        with torch.cuda.device(d.device):
            for n in range(self.n_iters + 1):
                _cabi.check(d.lib, d.lib.smo_kdyn_snapshot_coef(d.h, self.ptr(), self.n_iters, n, coef.data_ptr(), _stream_ptr()))
                out[..., n] = coef.view(3, d.nkx, d.Nc, d.Nc + 1)[c, :, :, :d.Nc].cpu().numpy()
        return out


class CheckpointStore(dict):
    """Two-level (revolve-style) checkpoint store for runs whose N_SUB_ITERS+1 states do not fit in HBM (BASELINE config 4:
    256^3 x 1000 steps = 401 GB).  The forward solve keeps the states 0, every, 2*every, ... and the final one; the adjoint
    sweep recomputes one segment of ``every`` states at a time.  ``rho`` = extra forward steps / N_ITERS (forward-recompute
    factor quoted next to every checkpointed number); results are bit-identical to the fully stored sweep."""

    def __init__(self, domain, n_iters, every):
        super().__init__()
        self.domain, self.n_iters, self.every = domain, int(n_iters), int(every)
        lib = domain.lib
        self.buf = torch.zeros(lib.smo_kdyn_checkpoint_bytes(domain.h, self.n_iters, self.every) // 16, dtype=torch.complex128,
                               device=domain.device)
        self.seg = torch.zeros(lib.smo_kdyn_segment_bytes(domain.h, self.every) // 16, dtype=torch.complex128, device=domain.device)
        self.valid = False
        nseg = (self.n_iters + self.every - 1) // self.every
        self.states_held = nseg + 1 + self.every + 1
        self.rho = sum(max(min(self.every, self.n_iters - k * self.every) - 1, 0) for k in range(nseg)) / max(self.n_iters, 1)   # full steps recomputed

    def ptr(self):
        return self.buf.data_ptr()


def GEN_BUFFER(Npts, domain, N_SUB_ITERS, checkpoint_every=None):
    """KD:319-355.  ``checkpoint_every``: None = keep every state in HBM if it fits (else sqrt(N)-spaced checkpoints),
    0 = always keep every state, k > 0 = checkpoint every k-th state (CheckpointStore)."""
    if checkpoint_every is None:
        need = domain.lib.smo_kdyn_snapshot_bytes(domain.h, int(N_SUB_ITERS))
        with torch.cuda.device(domain.device):
            free, _ = torch.cuda.mem_get_info()
        # every rank must take the same branch (the two stores issue different kernel / hand-shake sequences): MIN of "fits"
        fits = domain.all_agree(need < 0.8 * free, "min")
        checkpoint_every = 0 if fits else int(np.ceil(np.sqrt(max(int(N_SUB_ITERS), 1))))
    if checkpoint_every and checkpoint_every > 0:
        return CheckpointStore(domain, N_SUB_ITERS, checkpoint_every)
    return SnapshotStore(domain, N_SUB_ITERS)


# ---------------------------------------------------------------------------------------------------------------
def _as_slab(domain, x):
    if isinstance(x, DevVec):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    return domain.slab_from_host(x)


def _wrap_like(domain, proto, t):
    if isinstance(proto, DevVec):
        return DevVec(t)
    return domain.host_from_slab(t)


def Inner_Prod_3(x, y, domain, random_arg=None):
    """KD:173-181: (1/V) integ(x.y) dV = (1/M^3) * sum over all 3*M^3 entries of x_j*y_j (raw vectors)."""
    if isinstance(x, DevVec) or isinstance(y, DevVec) or domain.nranks == 1:
        xt, yt = _as_slab(domain, x), _as_slab(domain, y)
        v = domain.vecops(xt.numel()).dot(xt, yt, 1.0 / float(domain.M) ** 3)
        return domain.allreduce_sum(v)
    # replicated host vectors on several ranks: every rank reduces its slab, partials added in rank order
    xt, yt = domain.slab_from_host(x), domain.slab_from_host(y)
    v = domain.vecops(xt.numel()).dot(xt, yt, 1.0 / float(domain.M) ** 3)
    return domain.allreduce_sum(v)


def _flags(Cost_function, Adjoint_type):
    f = 0
    if Cost_function == "Integrated":
        f |= _cabi.SMO_COST_INTEGRATED
    elif Cost_function != "Final":
        raise ValueError(Cost_function)
    if Adjoint_type == "Continuous":
        f |= _cabi.SMO_ADJOINT_CONTINUOUS
    elif Adjoint_type != "Discrete":
        raise ValueError(Adjoint_type)
    return f


def FWD_Solve_IVP_Lin(X0, domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, Cost_function="Final", Adjoint_type="Discrete"):
    """KD:529-689.  Returns -J, J = <B^N,B^N> ("Final"); fills the snapshot store and caches U for the adjoint."""
    if N_SUB_ITERS != N_ITERS:
        raise NotImplementedError("N_SUB_ITERS != N_ITERS (the reference script sets them equal, KD:1031)")
    if N_ITERS != X_FWD_DICT.n_iters:
        raise ValueError("snapshot store was allocated for N_ITERS=%d" % X_FWD_DICT.n_iters)
    Bt, Ut = _as_slab(domain, X0[0]), _as_slab(domain, X0[1])
    J = C.c_double()
    with torch.cuda.device(domain.device):
        if isinstance(X_FWD_DICT, CheckpointStore):
            _cabi.check(domain.lib, domain.lib.smo_kdyn_forward_ckpt(domain.h, Bt.data_ptr(), Ut.data_ptr(), float(Rm), float(dt),
                                                                     int(N_ITERS), X_FWD_DICT.every, X_FWD_DICT.ptr(), C.byref(J),
                                                                     _flags(Cost_function, "Discrete"), _stream_ptr()))
        else:
            _cabi.check(domain.lib, domain.lib.smo_kdyn_forward(domain.h, Bt.data_ptr(), Ut.data_ptr(), float(Rm), float(dt),
                                                                int(N_ITERS), X_FWD_DICT.ptr(), C.byref(J),
                                                                _flags(Cost_function, "Discrete"), _stream_ptr()))
    X_FWD_DICT.valid = True
    X_FWD_DICT.tag = (fingerprint(X0[0]), fingerprint(X0[1]), float(Rm), float(dt), int(N_ITERS), Cost_function)
    if sideout.enabled(SIDE_OUTPUTS):
        _side_outputs(domain, Bt, Ut, dt, int(N_ITERS), X_FWD_DICT)
    return (-1.) * domain.allreduce_sum(J.value)


def _side_outputs(domain, Bt, Ut, dt, n_iters, store):
    """KD:606-613: CheckPoints (iterations 0 and N: A, B, C and the projected velocity on the 3/2 grid) and scalar_data
    ("Magnetic energy" every 20 iterations; a CheckpointStore only holds the states its spacing keeps)"""
    if isinstance(store, CheckpointStore):
        have = [n for n in range(0, n_iters + 1, 20) if n % store.every == 0 or n == n_iters]
        coef_of = lambda n: store.buf.view(-1, 3 * domain.csize)[n // store.every if n != n_iters else -1].clone()
    else:
        have = list(range(0, n_iters + 1, 20))

        def coef_of(n):
            c = torch.zeros(3 * domain.csize, dtype=torch.complex128, device=domain.device)
            with torch.cuda.device(domain.device):
                _cabi.check(domain.lib, domain.lib.smo_kdyn_snapshot_coef(domain.h, store.ptr(), n_iters, n, c.data_ptr(), _stream_ptr()))
            return c
    en = []
    for n in have:
        g = to_grid(domain, coef_of(n))
        en.append(Inner_Prod_3(DevVec(g), DevVec(g), domain))
    last = to_grid(domain, coef_of(n_iters))
    Up = to_grid(domain, to_coef(domain, Ut).reshape(-1))          # parameter fields are projected on the retained modes [D2-8]
    Bf, Bl, Uh = domain.host_from_slab(Bt), domain.host_from_slab(last), domain.host_from_slab(Up)
    if domain.rank == 0:
        sideout.kdyn_outputs(domain, have, en, Bf, Bl, Uh, dt, n_iters)


def ADJ_Solve_IVP_Lin(X0, domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, Cost_function="Final", Adjoint_type="Discrete"):
    """KD:766-1004.  Returns [dJ/dB0, dJ/dU] in the layout/type of X0."""
    if not X_FWD_DICT.valid:
        raise RuntimeError("ADJ_Solve_IVP_Lin needs the snapshots of a preceding FWD_Solve_IVP_Lin (KD:955-957)")
    stale = getattr(X_FWD_DICT, "tag", None) != (fingerprint(X0[0]), fingerprint(X0[1]), float(Rm), float(dt), int(N_ITERS), Cost_function)
    if domain.all_agree(stale, "max"):    # (sharded vectors: one rank's slab may differ while another's does not - MAX of "stale")
        # the store was written for another X or other parameters (never happens in the reference optimiser): refill it
        FWD_Solve_IVP_Lin(X0, domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, Cost_function, Adjoint_type)
    gB = torch.empty(3 * domain.gsize, dtype=torch.float64, device=domain.device)
    gU = torch.empty(3 * domain.gsize, dtype=torch.float64, device=domain.device)
    with torch.cuda.device(domain.device):
        if isinstance(X_FWD_DICT, CheckpointStore):
            _cabi.check(domain.lib, domain.lib.smo_kdyn_adjoint_ckpt(domain.h, float(Rm), float(dt), int(N_ITERS), X_FWD_DICT.every,
                                                                     X_FWD_DICT.ptr(), X_FWD_DICT.seg.data_ptr(), gB.data_ptr(),
                                                                     gU.data_ptr(), _flags(Cost_function, Adjoint_type), _stream_ptr()))
        else:
            _cabi.check(domain.lib, domain.lib.smo_kdyn_adjoint(domain.h, float(Rm), float(dt), int(N_ITERS), X_FWD_DICT.ptr(),
                                                                gB.data_ptr(), gU.data_ptr(),
                                                                _flags(Cost_function, Adjoint_type), _stream_ptr()))
    return [_wrap_like(domain, X0[0], gB), _wrap_like(domain, X0[1], gU)]


def FWD_Solve_IVP_Prep(Bx0, Ux0, domain, Rm, dt, N_ITERS):
    """KD:452-527: N_ITERS+1 CNAB1 steps; returns the final field on the grid as a device slab [3][M][M][nz]."""
    Bt, Ut = _as_slab(domain, Bx0), _as_slab(domain, Ux0)
    out = torch.empty_like(Bt)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_kdyn_prep(domain.h, Bt.data_ptr(), Ut.data_ptr(), float(Rm), float(dt),
                                                         int(N_ITERS), out.data_ptr(), _stream_ptr()))
    return out


def to_coef(domain, g):
    """grid slab [3][M][M][nz] -> coefficients [3][nkx][Nc][Nc+1] complex128 (last column is padding)"""
    gt = _as_slab(domain, g)
    c = torch.zeros(3 * domain.csize, dtype=torch.complex128, device=domain.device)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_kdyn_to_coef(domain.h, gt.data_ptr(), c.data_ptr(), _stream_ptr()))
    return c.view(3, domain.nkx, domain.Nc, domain.Nc + 1)


def to_grid(domain, c):
    g = torch.empty(3 * domain.gsize, dtype=torch.float64, device=domain.device)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_kdyn_to_grid(domain.h, c.contiguous().data_ptr(), g.data_ptr(), _stream_ptr()))
    return g


def _wavenumbers(domain):
    kf = 2.0 * np.pi / domain.L
    kx = kf * torch.arange(domain.rank * domain.nkx, (domain.rank + 1) * domain.nkx, dtype=torch.float64, device=domain.device)
    n = np.concatenate([np.arange(0, domain.kmax + 1), np.arange(-domain.kmax, 0), [0]])   # + padding column
    kc = kf * torch.from_numpy(n.astype(np.float64)).to(domain.device)
    return kx.view(-1, 1, 1), kc[:domain.Nc].view(1, -1, 1), kc.view(1, 1, -1)


def Generate_IC(Npts, X=(0., 2. * np.pi), M_0=1.0, U_Noise=False, Rm=1.0, dt=5e-04, device=None, as_devvec=False, seeds=(42, 42)):
    """KD:183-317.  ``Rm``/``dt`` stand for the module-level globals the reference's smoothing step uses instead of
    Rm_IC/dt_IC (KD:299-302 quirk); defaults are the literals at KD:1028-1029.  ``seeds`` = (B, U) noise seeds; the
    reference uses 42 for both (KD:224, 272)."""
    domain = Domain(Npts, X, device)
    M = domain.M
    kx, ky, kz = _wavenumbers(domain)

    def filt_mask():
        # KD:30-55: index/size > 0.25 on ANY axis of the (N/2, N-1, N-1) array is zeroed (index based)
        ix = torch.from_numpy(np.linspace(0, 1, domain.Nh, endpoint=False)[domain.rank * domain.nkx:(domain.rank + 1) * domain.nkx] <= 0.25)
        ic = np.linspace(0, 1, domain.Nc, endpoint=False) <= 0.25
        iy = torch.from_numpy(ic)
        iz = torch.from_numpy(np.concatenate([ic, [False]]))
        return (ix.view(-1, 1, 1) & iy.view(1, -1, 1) & iz.view(1, 1, -1)).to(domain.device)

    def curl_of_noise(seed):
        rand = np.random.RandomState(seed=seed)
        noise = rand.standard_normal((M, M, M))
        z = np.zeros_like(noise)
        phi = to_coef(domain, np.concatenate([noise.ravel(), z.ravel(), z.ravel()]))[0] * filt_mask()
        grads = torch.stack([1j * kx * phi, 1j * ky * phi, 1j * kz * phi])
        g = to_grid(domain, grads).view(3, -1)
        px, py, pz = g[0], g[1], g[2]
        return torch.cat([py - pz, pz - px, px - py])                     # KD:241-243

    B = curl_of_noise(seeds[0])
    if U_Noise is False:
        g = domain.interval[0] + domain.L * np.arange(M) / M
        x = g[:, None, None]; y = g[None, :, None]; z = g[None, None, domain.z0:domain.z0 + domain.nz]
        one = np.ones((M, M, domain.nz))
        U = np.concatenate([(0.5 * np.sin(y) * np.cos(z) / np.sqrt(3.) * one).ravel(),     # KD:258-260
                            (0.5 * np.sin(z) * np.cos(x) / np.sqrt(3.) * one).ravel(),
                            (0.5 * np.sin(x) * np.cos(y) / np.sqrt(3.) * one).ravel()])
        U = torch.from_numpy(U).to(domain.device)
    else:
        U = curl_of_noise(seeds[1])
    ip = lambda a: Inner_Prod_3(DevVec(a), DevVec(a), domain)
    U = U * np.sqrt(1. / ip(U))                                             # KD:287-292
    B = FWD_Solve_IVP_Prep(B, U, domain, Rm, dt, 100)                       # KD:296-302
    B = B * np.sqrt(M_0 / ip(B))                                            # KD:305-310
    if as_devvec:
        return domain, DevVec(B), DevVec(U)
    return domain, domain.host_from_slab(B), domain.host_from_slab(U)


def File_Manips(k):
    """KD:1006-1021: keeps scalar_data / CheckPoints of optimiser iteration k (needs SIDE_OUTPUTS; .npz, and .h5 with h5py)."""
    return sideout.file_manips(k)
