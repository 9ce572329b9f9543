"""B200 drop-in for Example_Problems/Periodic_Domain(Fourier)/Swift_Hohenberg/FWD_Solve_SH23.py (alias ``SH``).

Same function names, argument order and return conventions as the reference, so that its driver block (SH:750-784)
works unchanged with these callables:

    domain, X_0 = Generate_IC(E_0)                       # SH:174-236
    X_FWD_DICT  = GEN_BUFFER(domain, N_SUB_ITERS)        # SH:238-272
    args_IP = (domain, None)
    args_f  = [domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, None, "Discrete"]
    Optimise_On_Multi_Sphere([X_0], [E_0], FWD_Solve_IVP_Lin, ADJ_Solve_IVP_Lin, Inner_Prod, args_f, args_IP, ...)

``domain`` is a small stand-in for the dedalus domain (opaque to the optimiser); ``X_FWD_DICT`` is a device-resident
snapshot store.  Vectors are either float64 numpy arrays of 2*Npts dealiased-grid values (the reference's layout,
SH:110-128; "Mode H": copied to/from the GPU inside each call) or ``DevVec`` objects ("Mode D": resident in HBM).
All arithmetic happens in libsmo_b200.so (sh23.cuh: one kernel launch per solve); there is no CPU fallback.

Batched ensembles (BASELINE config 5): every function also accepts vectors holding ``batch`` instances back to back
(shape [batch*M]) when the domain was made with ``batch > 1``; J is then returned per instance by ``forward_batch``.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi, sideout
from .devvec import DevVec, VecOps, _stream_ptr, fingerprint

PARAM_A = -0.3   # SH:309
SIDE_OUTPUTS = False   # True (or SMO_SIDE_OUTPUTS=1): write the reference's CheckPoints / scalar_data handlers (sideout.py)


class Domain:
    """Stand-in for the dedalus domain of SH:202-204 (one Fourier basis, Npts modes, dealias 2)."""

    def __init__(self, Npts=256, X=(0., 12. * np.pi), device="cuda:0", a=PARAM_A):
        self.lib = _cabi.load()
        self.N = int(Npts)
        self.dealias = 2
        self.M = 2 * self.N
        self.Nh = self.N // 2
        self.interval = (float(X[0]), float(X[1]))
        self.L = self.interval[1] - self.interval[0]
        self.hypervolume = self.L
        self.device = torch.device(device)
        self.a = float(a)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib, self.lib.smo_sh23_create(C.byref(h), self.N, self.L, self.a))
        self.h = h
        self._vecops = {}

    def vecops(self, n):
        if n not in self._vecops:
            self._vecops[n] = VecOps(n, self.device)
        return self._vecops[n]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.smo_sh23_destroy(self.h)
                self.h = None
        except Exception:
            pass


class SnapshotStore(dict):
    """Device-resident replacement of the reference's ``{'A_fwd': complex128[Npts/2, N_SUB_ITERS+1]}`` (SH:266-272).

    Opaque HBM store: per instance the N_SUB_ITERS+1 forward states ON THE GRID (M doubles each: what the adjoint's pointwise
    product reads; written from the registers of the forward solve) + the coefficients of the final state.  ``['A_fwd']``
    converts back (one r2c launch per state) and returns a host copy in the reference's [Npts/2, N_SUB_ITERS+1] orientation."""

    def __init__(self, domain, n_iters, batch=1, buf=None):
        super().__init__()
        self.domain, self.n_iters, self.batch = domain, int(n_iters), int(batch)
        self.inst_bytes = domain.lib.smo_sh23_snapshot_bytes(domain.h, self.n_iters)
        n = self.inst_bytes * self.batch // 8
        if buf is None:
            buf = torch.zeros(n, dtype=torch.float64, device=domain.device)
        elif buf.dtype != torch.float64 or buf.dim() != 1 or buf.numel() != n or not buf.is_contiguous():
            raise ValueError("snapshot store over an existing buffer: need a contiguous float64 vector of %d entries" % n)
        self.buf = buf       # (instances are contiguous rows: the first k rows of a larger store are a k-instance store)
        self.valid = False

    def ptr(self):
        return self.buf.data_ptr()

    def coef(self, n):
        """coefficients [batch][Npts/2] of stored state n (device tensor)"""
        c = torch.empty(self.batch * self.domain.Nh, dtype=torch.complex128, device=self.domain.device)
        with torch.cuda.device(self.domain.device):
            _cabi.check(self.domain.lib, self.domain.lib.smo_sh23_snapshot_coef(self.domain.h, self.ptr(), self.batch, self.n_iters, int(n),
                                                                                c.data_ptr(), _stream_ptr()))
        return c.view(self.batch, self.domain.Nh)

    def __getitem__(self, key):
        if key != 'A_fwd':
            raise KeyError(key)
        a = torch.stack([self.coef(n) for n in range(self.n_iters + 1)], dim=2).cpu().numpy()     # [batch][Nh][n]
        return a[0].copy() if self.batch == 1 else a.copy()


def GEN_BUFFER(domain, N_SUB_ITERS, Npts=256, batch=1):
    """SH:238-272"""
    return SnapshotStore(domain, N_SUB_ITERS, batch)


# ---------------------------------------------------------------------------------------------------------------
def _as_dev(domain, x):
    """-> (float64 CUDA tensor, was_devvec)"""
    if isinstance(x, DevVec):
        return x.t, True
    if isinstance(x, torch.Tensor):
        return x, True
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return t.to(domain.device, non_blocking=True), False


def Inner_Prod(x, y, domain, rand_arg=None):
    """SH:158-172: (1/L) integ(x*y) dx = mean over the dealiased grid of x_j*y_j (raw vectors)."""
    xt, _ = _as_dev(domain, x)
    yt, _ = _as_dev(domain, y)
    n = xt.numel()
    return domain.vecops(n).dot(xt, yt, 1.0 / domain.M)


def Field_to_Vec(domain, Fx):
    """SH:89-128 - the field is already a flat grid vector here."""
    return Fx


def forward_batch(X, domain, dt, N_ITERS, X_FWD_DICT):
    """J[b] = dt*sum_n mean(u_n^2) for every instance of a batched vector (device tensor out)."""
    xt, _ = _as_dev(domain, X)
    batch = xt.numel() // domain.M
    if batch != X_FWD_DICT.batch or N_ITERS != X_FWD_DICT.n_iters:
        raise ValueError("snapshot store was allocated for batch=%d, N_ITERS=%d" % (X_FWD_DICT.batch, X_FWD_DICT.n_iters))
    J = torch.empty(batch, dtype=torch.float64, device=domain.device)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_sh23_forward(domain.h, xt.data_ptr(), batch, float(dt), int(N_ITERS),
                                                            X_FWD_DICT.ptr(), J.data_ptr(), _stream_ptr()))
    X_FWD_DICT.valid = True
    return J


def adjoint_batch(domain, dt, N_ITERS, X_FWD_DICT, Adjoint_type="Discrete"):
    if not X_FWD_DICT.valid:
        raise RuntimeError("ADJ_Solve_IVP_Lin needs the snapshots of a preceding FWD_Solve_IVP_Lin (SH:688)")
    batch = X_FWD_DICT.batch
    G = torch.empty(batch * domain.M, dtype=torch.float64, device=domain.device)
    flags = {"Discrete": 0, "Continuous": _cabi.SMO_ADJOINT_CONTINUOUS}[Adjoint_type]
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_sh23_adjoint(domain.h, batch, float(dt), int(N_ITERS), X_FWD_DICT.ptr(),
                                                            G.data_ptr(), flags, _stream_ptr()))
    return G


def FWD_Solve_IVP_Lin(X_k, domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, filename=None, Adjoint_type="Discrete"):
    """SH:409-545.  Returns -J, J = dt*sum_{n=0}^{N_ITERS} <u^n,u^n>; fills the snapshot store."""
    if N_SUB_ITERS != N_ITERS:
        raise NotImplementedError("N_SUB_ITERS != N_ITERS (the reference scripts always set them equal, SH:755)")
    if filename is not None:
        raise NotImplementedError("restart from a dedalus checkpoint file (SH:459-460)")
    J = forward_batch(X_k[0], domain, dt, N_ITERS, X_FWD_DICT)
    X_FWD_DICT.tag = (fingerprint(X_k[0]), float(dt), int(N_ITERS))
    Jh = J.cpu().numpy()
    if Jh.size == 1 and sideout.enabled(SIDE_OUTPUTS):      # SH:478-483 (one problem only, like the reference)
        sideout.sh23_outputs(domain, X_FWD_DICT['A_fwd'], dt, N_ITERS)
    return (-1.) * float(Jh[0]) if Jh.size == 1 else (-1.) * Jh


def ADJ_Solve_IVP_Lin(X_k, domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, filename=None, Adjoint_type="Discrete"):
    """SH:598-729.  Returns [dJ/du0] in the layout/type of X_k[0]."""
    if X_FWD_DICT.valid and getattr(X_FWD_DICT, "tag", None) != (fingerprint(X_k[0]), float(dt), int(N_ITERS)):
        # the store was written for another X (never happens in the reference optimiser): refill it
        FWD_Solve_IVP_Lin(X_k, domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, filename, Adjoint_type)
    G = adjoint_batch(domain, dt, N_ITERS, X_FWD_DICT, Adjoint_type)
    if isinstance(X_k[0], DevVec):
        return [DevVec(G)]
    out = torch.empty(G.numel(), dtype=G.dtype, pin_memory=True)    # page-locked D2H (see kdyn.Domain.host_from_slab)
    out.copy_(G)
    return [out.numpy()]


def FWD_Solve_IVP_PREP(X_k, domain, dt=1e-02, N_ITERS=100, N_SUB_ITERS=100):
    """SH:334-407: N_ITERS+1 SBDF1 steps; returns the final state on the dealiased grid (device tensor)."""
    xt, _ = _as_dev(domain, X_k)
    batch = xt.numel() // domain.M
    out = torch.empty_like(xt)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_sh23_prep(domain.h, xt.data_ptr(), batch, float(dt), int(N_ITERS),
                                                         out.data_ptr(), _stream_ptr()))
    return out


def to_coef(domain, x):
    xt, _ = _as_dev(domain, x)
    batch = xt.numel() // domain.M
    c = torch.empty(batch * domain.Nh, dtype=torch.complex128, device=domain.device)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_sh23_to_coef(domain.h, xt.data_ptr(), batch, c.data_ptr(), _stream_ptr()))
    return c


def to_grid(domain, c):
    batch = c.numel() // domain.Nh
    g = torch.empty(batch * domain.M, dtype=torch.float64, device=domain.device)
    with torch.cuda.device(domain.device):
        _cabi.check(domain.lib, domain.lib.smo_sh23_to_grid(domain.h, c.data_ptr(), batch, g.data_ptr(), _stream_ptr()))
    return g


def Generate_IC(E_0=1.0, Npts=256, X=(0., 12. * np.pi), device="cuda:0", as_devvec=False):
    """SH:174-236: seed-42 noise -> low-pass (index/size > 0.5 zeroed, SH:28-53) -> <u,u> = E_0 -> 101 SBDF1 steps at
    dt = 0.01 (SH:228) -> <u,u> = E_0."""
    domain = Domain(Npts, X, device)
    rand = np.random.RandomState(seed=42)
    noise = rand.standard_normal((domain.M,))
    c = to_coef(domain, noise)
    keep = torch.from_numpy(np.linspace(0, 1, domain.Nh, endpoint=False) <= 0.5).to(domain.device)
    phi = to_grid(domain, c * keep)
    ops = domain.vecops(domain.M)
    phi = phi * np.sqrt(E_0 / ops.dot(phi, phi, 1.0 / domain.M))
    phi = FWD_Solve_IVP_PREP(phi, domain)
    phi = phi * np.sqrt(E_0 / ops.dot(phi, phi, 1.0 / domain.M))
    return domain, (DevVec(phi) if as_devvec else phi.cpu().numpy())


def File_Manips(k):
    """SH:731-746: keeps scalar_data / CheckPoints of optimiser iteration k (needs SIDE_OUTPUTS; .npz, and .h5 with h5py)."""
    return sideout.file_manips(k)
