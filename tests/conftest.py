import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFSTUBS = os.path.join(ROOT, "tests", "refstubs")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def reference_dir():
    """where the UNMODIFIED reference optimiser files can be imported from: $SMO_REFERENCE_DIR, /root/reference (build
    container) or baseline/_ref (staged by tools/stage_reference.py; travels to the GPU box, git-ignored)"""
    for d in (os.environ.get("SMO_REFERENCE_DIR"), REFERENCE, os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "Sphere_Grad_Descent.py")) and os.path.isfile(os.path.join(d, "TestGrad.py")):
            return d
    return None


@pytest.fixture(scope="session")
def refopt():
    """The UNMODIFIED reference optimiser / gradient test.  Returns (Sphere_Grad_Descent, TestGrad) or skips."""
    d = reference_dir()
    if d is None:
        pytest.skip("reference optimiser files not present (no /root/reference, no baseline/_ref)")
    for p in (REFSTUBS, d):
        if p not in sys.path:
            sys.path.insert(0, p)
    import Sphere_Grad_Descent as SGD
    import TestGrad as TG
    return SGD, TG
