"""Build recipe of libsmo_b200.so (hand-written CUDA for sm_100a, in-tree so the .so travels with the snapshot).

    python -m spheremanopt_b200._build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libsmo_b200.so")
SOURCES = ["smo_api.cu"]
HEADERS = ["smo_common.cuh", "codelets.cuh", "fft_core.cuh", "fft_pass.cuh", "xpass.cuh", "xpass_half.cuh", "kd_epilogue.cuh", "zstep.cuh",
           "sh23.cuh", "reduce.cuh", os.path.join("..", "..", "include", "smo_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def nvcc_command(out=OUT, extra=()):
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-Xlinker", "-Bsymbolic", "-DSMO_WITH_NCCL", "-o", out]
    cmd += list(extra)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-ldl"]   # NCCL is dlopen'ed at run time (shares the host program's libnccl.so.2)
    return cmd


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    cmd = nvcc_command(extra=["-Xptxas", "-v"] if verbose else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
