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


@pytest.mark.gpu
def test_kdyn_two_gpus_staged_bulk_push():
    """SMO_OPT_BULK_PUSH = 3: the transposes pushed by the fused z step and the forward y pass go through shared-memory staging and
    TMA bulk stores into the peer's memory; same oracle parity (the 128^3 grid is the one whose z step has the staged path)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, SMO_KDYN_OPTS="11=3", MP_CASES="32:6,128:2", MP_VARIANTS="3,2,10")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29539", os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "MP_PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_sh23_ensemble_sharded_over_two_gpus():
    """BASELINE config 5's sharding on hardware: bench.py --workload sh23ens under torchrun with 2 ranks (2048 instances each, no
    collective on the data path) prints one JSON line whose per-instance cost matches the single-GPU kernels' scale"""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "sh23ens", "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-2000:] + r.stderr[-2000:]
    b = json.loads(lines[0])
    assert b["n_gpus"] == 2 and b["config"]["instances"] == 4096 and b["gpu_launches"] >= 4 and 0 < b["ms_per_step"] < 100
