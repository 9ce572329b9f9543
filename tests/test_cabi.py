"""The C-ABI library loads on a CPU-only box and exports exactly what include/smo_b200.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "smo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smo_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from spheremanopt_b200 import _build, _cabi
    _build.build()
    return _cabi.load()


def test_header_and_binding_agree():
    from spheremanopt_b200 import _cabi
    assert header_functions() == sorted(_cabi.SIGNATURES)


def test_header_constants_and_binding_agree():
    """every SMO_* integer constant of the header (flags, option keys) has the same value in the ctypes binding"""
    from spheremanopt_b200 import _cabi
    src = open(os.path.join(ROOT, "include", "smo_b200.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"^#define\s+(SMO_[A-Z0-9_]+)\s+(-?\d+)\b", src, flags=re.M)}
    assert {"SMO_ADJOINT_CONTINUOUS", "SMO_COST_INTEGRATED", "SMO_OPT_PDL", "SMO_OPT_BULK_PUSH"} <= set(consts)
    for k, v in consts.items():
        if hasattr(_cabi, k):
            assert getattr(_cabi, k) == v, k
    opts = {k for k in consts if k.startswith("SMO_OPT_") and k != "SMO_OPT_FUSED_Z"}
    assert opts <= set(dir(_cabi)), opts - set(dir(_cabi))
    assert len({consts[k] for k in opts}) == len(opts)      # option keys are distinct


def test_library_exports_every_declared_symbol(lib):
    from spheremanopt_b200 import _cabi
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (smo_[a-z0-9_]+)", out))
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing
    assert lib.smo_version() >= 100
    assert lib.smo_launch_count() == 0


def test_library_is_sm100a_cuda():
    from spheremanopt_b200 import _cabi
    r = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout


def test_every_phase_kernel_waits_for_its_grid_dependency():
    """programmatic dependent launches (SMO_OPT_PDL) are only correct if EVERY kernel a time loop can launch blocks in
    griddepcontrol.wait before it touches its predecessor's results: static check of the SASS (ACQBULK = griddepcontrol.wait,
    PREEXIT = griddepcontrol.launch_dependents) of every smo_kernel<> instantiation"""
    from spheremanopt_b200 import _cabi
    r = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    funcs = re.split(r"\n\s*Function : ", r.stdout)[1:]
    kernels = [f for f in funcs if "smo_kernel" in f.split("\n", 1)[0]]
    assert len(kernels) > 50
    for f in kernels:
        name = f.split("\n", 1)[0]
        assert "ACQBULK" in f and "PREEXIT" in f, name
        # the wait comes before the first global store / bulk store / atomic of the kernel
        first = {op: f.find(op) for op in ("ACQBULK", " STG", "UBLKCP.G.S", " ATOMG", " RED.")}
        for op, pos in first.items():
            assert op == "ACQBULK" or pos < 0 or pos > first["ACQBULK"], (name, op)


def test_argument_errors_without_gpu(lib):
    import ctypes as C
    h = C.c_void_p()
    rc = lib.smo_sh23_create(C.byref(h), 100, 37.7, -0.3)
    assert rc != 0 and b"not supported" in lib.smo_last_error()
    rc = lib.smo_kdyn_create(C.byref(h), 128, 6.28, 0, 3, None)
    assert rc != 0 and b"divide" in lib.smo_last_error()


def test_product_has_no_cpu_fallback(monkeypatch):
    """a missing CUDA library must raise, not fall back"""
    from spheremanopt_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libsmo_b200.so")
    with pytest.raises(ImportError):
        _cabi.load()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spheremanopt_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f
