"""Dedalus-v2 Fourier mode layout and transforms, restated with scipy.fft.  TEST INFRASTRUCTURE ONLY.

Restates the behaviour the reference relies on through ``de.Fourier`` /
``de.Domain`` (FWD_Solve_SH23.py:203-204, FWD_Solve_KDyn.py:213-216) and every
``field['g']`` / ``field['c']`` access.  The Dedalus sources are not under
/root/reference; the facts below are the [D2-n] items of SURVEY.md section 8(c):

[D2-1] first Fourier axis with a real grid dtype is r2c and keeps k = 0..N/2-1
       (Nyquist dropped); later Fourier axes are c2c and keep N-1 modes ordered
       [0..kmax, -kmax..-1], kmax = (N-1)//2.
[D2-2] grid->coeff divides by the grid size (coefficients are amplitudes of e^{ikx}).
[D2-3] a change of scales is zero-padding / truncation in coefficient space.
[D2-5] integ over a Fourier axis = L * c_0, so (1/V) integ(F) = grid mean of F.
"""
import numpy as np
import scipy.fft as sfft

WORKERS = 1  # bench.py raises this for the timed CPU baseline


class Domain:
    """Stand-in for the ``dedalus`` domain object that the reference passes around opaquely.

    Mirrors what Generate_IC builds (FWD_Solve_SH23.py:202-204: one Fourier axis,
    dealias 2; FWD_Solve_KDyn.py:212-216: three Fourier axes, dealias 3/2).
    """

    def __init__(self, Npts, interval, dealias, dim):
        self.N = int(Npts)
        self.dim = int(dim)
        self.interval = (float(interval[0]), float(interval[1]))
        self.L = self.interval[1] - self.interval[0]
        self.dealias = dealias
        M = self.N * dealias
        assert abs(M - round(M)) < 1e-12, "dealiased grid size must be an integer"
        self.M = int(round(M))
        self.Nh = self.N // 2            # retained r2c modes  [D2-1]
        self.kmax = (self.N - 1) // 2    # retained c2c modes are -kmax..kmax  [D2-1]
        self.Nc = 2 * self.kmax + 1
        self.hypervolume = self.L ** self.dim
        self.grid_shape = (self.M,) * self.dim
        self.coeff_shape = (self.Nh,) + (self.Nc,) * (self.dim - 1)
        self.vec_len = self.M ** self.dim

    # wavenumbers -------------------------------------------------------------------------
    def kx(self):
        return (2.0 * np.pi / self.L) * np.arange(self.Nh)

    def kc(self):
        n = np.concatenate([np.arange(0, self.kmax + 1), np.arange(-self.kmax, 0)])
        return (2.0 * np.pi / self.L) * n

    def csel(self):
        """indices of the retained c2c modes inside a length-M FFT output"""
        return np.concatenate([np.arange(0, self.kmax + 1), np.arange(self.M - self.kmax, self.M)])

    # 1-D transforms ----------------------------------------------------------------------
    def to_coef_1d(self, g):
        c = sfft.rfft(np.asarray(g, dtype=np.float64)) / self.M   # [D2-2]
        return c[: self.Nh].copy()                                # [D2-3]

    def to_grid_1d(self, c):
        full = np.zeros(self.M // 2 + 1, dtype=np.complex128)
        full[: self.Nh] = c
        return sfft.irfft(full, n=self.M) * self.M

    # 3-D transforms (x r2c, then y, then z going grid->coeff) -----------------------------
    def to_coef_3d(self, g):
        M = self.M
        sel = self.csel()
        a = sfft.rfft(np.asarray(g, dtype=np.float64).reshape(M, M, M), axis=0, workers=WORKERS)[: self.Nh]
        a = sfft.fft(a, axis=1, workers=WORKERS)[:, sel, :]
        a = sfft.fft(a, axis=2, workers=WORKERS)[:, :, sel]
        return a / float(M) ** 3

    def to_grid_3d(self, c):
        M = self.M
        sel = self.csel()
        a = np.zeros((self.Nh, self.Nc, M), dtype=np.complex128)
        a[:, :, sel] = c
        a = sfft.ifft(a, axis=2, workers=WORKERS)
        b = np.zeros((self.Nh, M, M), dtype=np.complex128)
        b[:, sel, :] = a
        b = sfft.ifft(b, axis=1, workers=WORKERS)
        full = np.zeros((M // 2 + 1, M, M), dtype=np.complex128)
        full[: self.Nh] = b
        # c2r discards Im of the kx=0 line, exactly as FFTW's c2r does
        return sfft.irfft(full, n=M, axis=0, workers=WORKERS) * float(M) ** 3


def domain_sh23(Npts=256, X=(0.0, 12.0 * np.pi)):
    """FWD_Solve_SH23.py:202-204"""
    return Domain(Npts, X, 2, 1)


def domain_kdyn(Npts=24, X=(0.0, 2.0 * np.pi)):
    """FWD_Solve_KDyn.py:212-216"""
    return Domain(Npts, X, 1.5, 3)
