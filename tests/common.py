"""shared helpers of the test-suite (seeded band-limited inputs)"""
import numpy as np

from oracle import kdyn as okd
from oracle import sh23 as osh


def sh23_input(dom, seed=1, amp=0.05, decay=0.1):
    rng = np.random.RandomState(seed)
    c = (rng.standard_normal(dom.Nh) + 1j * rng.standard_normal(dom.Nh)) * np.exp(-decay * np.arange(dom.Nh))
    c[0] = c[0].real
    return dom.to_grid_1d(c) * amp


def kdyn_field(dom, seed, decay=0.3, solenoidal=True):
    """band-limited (optionally solenoidal, zero-mean) vector field on the dealiased grid, unit norm"""
    M = dom.M
    r = np.random.RandomState(seed)
    c = [dom.to_coef_3d(r.standard_normal((M, M, M))) for _ in range(3)]
    K = okd._K(dom)
    c = [ci * np.exp(-decay * np.sqrt(K[3])) for ci in c]
    if solenoidal:
        c = okd._project(K, c)
    v = okd.Field_to_Vec(dom, *[dom.to_grid_3d(ci) for ci in c])
    return v / np.sqrt(okd.Inner_Prod_3(v, v, dom))


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
