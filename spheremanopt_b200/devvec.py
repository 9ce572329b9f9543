"""Device-resident vectors for the unmodified reference optimiser ("Mode D" of SURVEY.md section 8(b)).

``Optimise_On_Multi_Sphere`` (Sphere_Grad_Descent.py:692-838) and ``Adjoint_Gradient_Test`` (TestGrad.py:5-156)
only ever do ``v + v``, ``v - v``, ``scalar * v``, ``v * scalar``, unary minus and ``copy.deepcopy`` on the vectors
they are handed (SGD:642, 659, 687-690, 734, 772, 776; TG:89).  ``DevVec`` implements exactly that on a float64
buffer in HBM, each operation being one launch of the axpby kernel of libsmo_b200 (smo_vec_axpby), so an
optimisation keeps its iterates on the GPU and only scalars cross PCIe.

The class deliberately defines no ``__len__`` / ``__getitem__`` / ``__array__`` and sets ``__array_ufunc__ = None``:
``np.atleast_1d([...])`` (SGD:111) must see an opaque object, and ``np.float64 * DevVec`` (SGD:734) must defer to
``DevVec.__rmul__``.
"""
import numbers

import numpy as np
import torch

from . import _cabi


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class DevVec:
    __array_ufunc__ = None
    __slots__ = ("t",)

    def __init__(self, tensor):
        if tensor.dtype != torch.float64 or not tensor.is_cuda or not tensor.is_contiguous():
            raise TypeError("DevVec needs a contiguous float64 CUDA tensor")
        self.t = tensor

    # construction / export ----------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, device="cuda"):
        return cls(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device))

    def numpy(self):
        return self.t.cpu().numpy()

    @property
    def n(self):
        return self.t.numel()

    def ptr(self):
        return self.t.data_ptr()

    # algebra --------------------------------------------------------------------------------
    def _axpby(self, a, b, other):
        lib = _cabi.load()
        out = torch.empty_like(self.t)
        with torch.cuda.device(self.t.device):
            _cabi.check(lib, lib.smo_vec_axpby(float(a), self.ptr(), float(b), other.ptr() if other is not None else None,
                                               out.data_ptr(), self.n, _stream_ptr()))
        return DevVec(out)

    def __add__(self, o):
        if isinstance(o, DevVec):
            return self._axpby(1.0, 1.0, o)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, DevVec):
            return self._axpby(1.0, -1.0, o)
        return NotImplemented

    def __rsub__(self, o):
        if isinstance(o, DevVec):
            return o._axpby(1.0, -1.0, self)
        return NotImplemented

    def __mul__(self, s):
        if isinstance(s, (numbers.Real, np.floating, np.integer)) or (isinstance(s, np.ndarray) and s.ndim == 0):
            return self._axpby(float(s), 0.0, None)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, s):
        if isinstance(s, (numbers.Real, np.floating, np.integer)):
            return self._axpby(1.0 / float(s), 0.0, None)
        return NotImplemented

    def __neg__(self):
        return self._axpby(-1.0, 0.0, None)

    def __pos__(self):
        return self

    def __deepcopy__(self, memo):
        return DevVec(self.t.clone())

    __copy__ = lambda self: DevVec(self.t.clone())

    def __repr__(self):
        return "DevVec(n=%d, device=%s)" % (self.n, self.t.device)


class VecOps:
    """dot / project / retract kernels on one device with a reusable reduction workspace (rows A5/B5, C1-C3)."""

    def __init__(self, n, device):
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.work = torch.empty(self.lib.smo_vec_work_bytes(self.n), dtype=torch.uint8, device=self.device)

    def dot(self, x, y, scale):
        import ctypes as C
        out = C.c_double()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib, self.lib.smo_vec_dot(x.data_ptr(), y.data_ptr(), self.n, float(scale), C.byref(out),
                                                       self.work.data_ptr(), _stream_ptr()))
        return out.value

    def project(self, x, v):
        """v - (<x,v>/<x,x>) x   (tangent_vector / transport_vector, SGD:625-659) without a host round trip"""
        out = torch.empty_like(v)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib, self.lib.smo_vec_project(x.data_ptr(), v.data_ptr(), out.data_ptr(), self.n,
                                                           self.work.data_ptr(), _stream_ptr()))
        return out

    def retract(self, x, alpha, d, M0, scale):
        """(x + alpha d) * sqrt(M0 / (scale * |x + alpha d|^2))   (Update_vector, SGD:661-690)"""
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib, self.lib.smo_vec_retract(x.data_ptr(), float(alpha), d.data_ptr(), float(M0), float(scale),
                                                           out.data_ptr(), self.n, self.work.data_ptr(), _stream_ptr()))
        return out


_HASH_WORK = {}


def fingerprint(x):
    """identity check of a vector: lets Grad_f notice that the snapshot store it is about to replay was written for a
    different X (the reference couples f and Grad_f through that store).
    Device vectors: (length, 64-bit position-sensitive checksum of every entry) - one streaming pass of the library's
    checksum kernel (~30 us for 170 MB), any changed entry changes it.  Host vectors: the same idea on the host, over every
    entry up to 2^20 entries and over a 65536-entry strided sample above that (a full host pass over a 170 MB vector would
    cost more than the H2D copy of the call itself; the optimiser only ever hands over vectors that differ everywhere)."""
    if isinstance(x, DevVec):
        x = x.t
    if isinstance(x, torch.Tensor):
        import ctypes as C
        lib = _cabi.load()
        n = x.numel()
        key = (x.device, n)
        if key not in _HASH_WORK:
            _HASH_WORK[key] = torch.empty(lib.smo_vec_work_bytes(n), dtype=torch.uint8, device=x.device)
        out = C.c_ulonglong()
        with torch.cuda.device(x.device):
            _cabi.check(lib, lib.smo_vec_checksum(x.data_ptr(), n, C.byref(out), _HASH_WORK[key].data_ptr(), _stream_ptr()))
        return (n, int(out.value))
    a = np.asarray(x).reshape(-1)
    n = a.size
    if n > (1 << 20):
        a = a[::max(1, n // 65536)]
    b = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    w = (2 * np.arange(b.size, dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        return (n, int((b * w).sum(dtype=np.uint64)))
