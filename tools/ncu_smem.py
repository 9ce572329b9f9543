"""top shared-memory instructions by wavefronts / excess: python tools/ncu_smem.py rep kernel-substring"""
import csv, io, subprocess, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; hdr = None; rows = []
def flush():
    if kern and sub in kern and rows:
        print("##", kern[:110])
        tot_w = sum(r[2] for r in rows); tot_x = sum(r[3] for r in rows); tot_i = sum(r[4] for r in rows)
        print("   total smem wavefronts %d, excessive %d, ideal %d ; LDS/STS/LDGSTS warp-instr %d" % (tot_w, tot_x, tot_w - tot_x, tot_i))
        agg = {}
        for r in rows:
            op = r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]
            a = agg.setdefault(op, [0, 0, 0]); a[0] += r[2]; a[1] += r[3]; a[2] += r[4]
        for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("   %-22s wavefronts %10d  excessive %10d  instr %9d  (%.2f wf/instr)" % (op, a[0], a[1], a[2], a[0] / max(a[2], 1)))
        for r in sorted(rows, key=lambda r: -r[3])[:12]:
            print("      excess %8d  wf %8d  n %7d  %s" % (r[3], r[2], r[4], r[1][:90]))
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "Kernel Name":
        flush(); kern = row[1]; rows = []; hdr = None; continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None: continue
    w = int(row[hdr["L1 Wavefronts Shared"]] or 0)
    if w == 0: continue
    x = int(row[hdr["L1 Wavefronts Shared Excessive"]] or 0)
    rows.append((row[hdr["Address"]], row[hdr["Source"]].strip(), w, x, int(row[hdr["Instructions Executed"]] or 0)))
flush()
