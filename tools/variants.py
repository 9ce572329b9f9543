"""development tool: build tuning variants of the library (-D knobs) into build/variants/ for A/B timing on the GPU box

    python tools/variants.py name1:-DSMO_PASS_MB=2,-DSMO_TY=16 name2:...
"""
import os
import subprocess
import sys

sys.path.insert(0, ".")
from spheremanopt_b200 import _build

os.makedirs("build/variants", exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = "build/variants/libsmo_%s.so" % name
    cmd = _build.nvcc_command(out=out, extra=[f for f in flags.split(",") if f])
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + o[-2000:])
