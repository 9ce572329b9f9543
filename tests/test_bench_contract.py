"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the driver's keys, and the GPU
arm refuses to run without a CUDA device (no CPU fallback in the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--workload", "kdyn24", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in b, key
    assert b["impl"] == "reference" and b["dtype"] == "f64" and b["value"] > 0 and b["vs_baseline"] is None
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1
    assert b["e2e"]["h2d_bytes_per_step"] == 0 and b["e2e"]["d2h_bytes_per_step"] == 0 and b["e2e"]["value"] == b["value"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--workload", "kdyn24", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--workload", "kdyn24", "--steps", "1", "--warmup", "1", "--no-cpu"])
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)
