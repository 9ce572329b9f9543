"""Stage the UNMODIFIED reference optimiser files next to the repo for the GPU box (which has no /root/reference).

    python tools/stage_reference.py

The reference has no setup.py / pyproject (``pip install --target baseline/_ref /root/reference`` has nothing to install), so its
"install" is a verbatim copy of its two top-level modules - Sphere_Grad_Descent.py (Optimise_On_Multi_Sphere, SGD:692) and
TestGrad.py (Adjoint_Gradient_Test, TG:5) - into ``baseline/_ref/`` (git-ignored: never part of the history; NOT
gpurun-ignored: it travels with the snapshot).  tests/test_gpu_history.py and bench.py's optimiser workloads import them
from there; nothing of the product imports them.  Dedalus-dependent example scripts are not staged (they cannot run).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SMO_REFERENCE_DIR", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("Sphere_Grad_Descent.py", "TestGrad.py")


def stage():
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not present at %s" % SRC, DST)
    sys.exit(0)
