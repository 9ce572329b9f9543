"""Stub so the UNMODIFIED reference Sphere_Grad_Descent.py imports without HDF5 (SGD:6).  The only
use is inside a bare try/except (SGD:821-829), so raising here reproduces 'h5py not available'.
Test infrastructure only."""


def File(*a, **k):
    raise OSError("h5py stub: no HDF5 in this image")
