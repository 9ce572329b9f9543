"""per-kernel SASS opcode histogram weighted by executed instructions: python tools/ncu_sass.py rep [kernel-substring]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; hdr = None; hist = None; tot = 0
def flush():
    if kern and hist and sub in kern:
        print("##", kern[:120], " total warp-instr:", tot)
        for op, n in hist.most_common(22):
            print("   %-14s %10d  %5.1f%%" % (op, n, 100.0 * n / tot))
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "Kernel Name":
        flush(); kern = row[1]; hist = collections.Counter(); tot = 0; hdr = None; continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None: continue
    src = row[hdr["Source"]].strip()
    n = int(row[hdr["Instructions Executed"]] or 0)
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0]
    hist[op] += n; tot += n
flush()
