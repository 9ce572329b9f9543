"""pytest wrapper of tests/mp_parity.py: slab-decomposed dynamo on 2 GPUs of one box against the oracle (all transports).
Skipped on boxes with a single GPU (the CPU suite covers the same host logic with gloo + emulated kernels)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_kdyn_two_gpus_all_transports():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MP_PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
