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


@pytest.fixture(scope="session")
def refopt():
    """The UNMODIFIED reference optimiser / gradient test, imported from /root/reference when it is there
    (build container only).  Returns (Sphere_Grad_Descent, TestGrad) or skips."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("/root/reference not present on this box")
    for p in (REFSTUBS, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import Sphere_Grad_Descent as SGD
    import TestGrad as TG
    return SGD, TG
