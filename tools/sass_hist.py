"""static SASS opcode histogram of the hot kernels of libsmo_b200.so (cuobjdump -sass; no GPU needed):

    python tools/sass_hist.py > profiles/r2_sass_hist.txt

Shows, per kernel, the opcode counts that matter for the review: fp64 arithmetic (DFMA/DADD/DMUL), shared-memory traffic
(LDS/STS), the staging engines (LDGSTS = 16-byte cp.async; UBLKCP = TMA bulk copy cp.async.bulk; UTMALDG = TMA tensor copy
cp.async.bulk.tensor; SYNCS = mbarrier), global
accesses, barriers and shuffles."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spheremanopt_b200", "libsmo_b200.so")
WANT = [("adjoint x pass 128^3 (XFused<Fac<16,12>, X_ADJ, GACC>)", "6XFusedINS_3FacILi16ELi12EEELi3ELb0ELb1EEE"),
        ("forward x pass 128^3 (XFused<Fac<16,12>, X_FWD>)", "6XFusedINS_3FacILi16ELi12EEELi2ELb0ELb0EEE"),
        ("adjoint x pass 256^3 (XFusedH<Fac<16,12>, X_ADJ, GACC>)", "7XFusedHINS_3FacILi16ELi12EEELi3ELb0ELb1ELi2EEE"),
        ("fused z step 128^3 (ZStep<Fac<16,12>,2>)", "5ZStepINS_3FacILi16ELi12EEELi2EEE"),
        ("y pass 128^3 inverse (FftPass<Fac<16,12>,+1,true,8>)", "7FftPassINS_3FacILi16ELi12EEELi1ELb1ELi8ELb0EEE"),
        ("y pass 128^3 forward, staged bulk push (FftPass<Fac<16,12>,-1,true,8,STAGE>)", "7FftPassINS_3FacILi16ELi12EEELin1ELb1ELi8ELb1EEE"),
        ("SH23 adjoint, ensembles (Sh23Adj<Fac<16,16>,4,8>)", "7Sh23AdjINS_3FacILi16ELi16EEELi4ELi8EEE"),
        ("Inner_Product (VecKernel<V_DOT>)", "9VecKernelILi0EEE")]
KEYS = ["DFMA", "DADD", "DMUL", "LDS", "STS", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDG", "STG", "CCTL", "BAR", "WARPSYNC", "SHFL", "MEMBAR", "ATOMG", "RED",
        "ACQBULK", "PREEXIT"]     # (ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents: programmatic dependent launch)

names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", names)
print("static SASS opcode counts, %s" % os.path.relpath(LIB, ROOT))
for title, sub in WANT:
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if sub not in name:
            continue
        hist = collections.Counter()
        for ln in f.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", ln)
            if m:
                hist[m.group(1)] += 1
        tot = sum(hist.values())
        print("\n## %s\n   %s\n   total %d instructions" % (title, name[:110], tot))
        print("   " + "  ".join("%s %d" % (k, hist.get(k, 0)) for k in KEYS))
        break
    else:
        print("\n## %s: not found" % title, file=sys.stderr)
