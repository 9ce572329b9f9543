"""Host emulation of the CUDA kernel bodies.  TEST INFRASTRUCTURE ONLY - never imported by the product package.

The kernels of spheremanopt_b200/csrc are written as barrier-separated phases (smo_common.cuh).  With
-DSMO_EMUL the same phase bodies are compiled by g++ and executed thread by thread on the host, which lets the
CPU test-suite (no GPU in the build container) check the index logic of every kernel against the oracle.
"Device" pointers of this build are plain host pointers, so numpy arrays are passed directly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from spheremanopt_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "spheremanopt_b200", "csrc")
_OUT = os.path.join(_HERE, "_build", "libsmo_emul.so")

_lib = None
_variants = {}


def _stale():
    if not os.path.exists(_OUT):
        return True
    t = os.path.getmtime(_OUT)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_ROOT, "include", "smo_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            cmd = ["g++", "-O2", "-std=c++17", "-x", "c++", "-DSMO_EMUL", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", _OUT,
                   os.path.join(_CSRC, "smo_api.cu")]
            subprocess.run(cmd, check=True)
        _lib = _cabi.bind(C.CDLL(_OUT))
    return _lib


def lib_variant(name, defines):
    """a second build of the emulation library with extra -D switches (e.g. the test-only radix-24 factorisation of M = 48)"""
    if name not in _variants:
        out = os.path.join(_HERE, "_build", "libsmo_emul_%s.so" % name)
        deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_ROOT, "include", "smo_b200.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            cmd = ["g++", "-O2", "-std=c++17", "-x", "c++", "-DSMO_EMUL"] + ["-D" + d for d in defines] + \
                  ["-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", out, os.path.join(_CSRC, "smo_api.cu")]
            subprocess.run(cmd, check=True)
        _variants[name] = _cabi.bind(C.CDLL(out))
    return _variants[name]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def check(rc):
    _cabi.check(lib(), rc)
