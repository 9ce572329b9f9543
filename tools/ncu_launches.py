"""summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of the bench command into per-kernel shares:

    python tools/ncu_launches.py gpurun_out/launches.csv "title" > profiles/rNx_launches_bench_kdyn128.md

The steady-state part of the forward and of the adjoint time loop is found by the 4-launch pattern y-inverse, fused x,
y-forward, fused z step (X_FWD = mode 2, X_ADJ = mode 3 in the kernel names)."""
import collections
import csv
import re
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].replace("void smo_kernel<", "").replace(">(Params)", "").replace("smo::", "")
    rows.append((name, float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
print("# %s" % title)
print("\n%d launches in the list, %.1f us of kernel time\n" % (len(rows), sum(r[1] for r in rows)))


def table(sel, what):
    tot = sum(r[1] for r in sel)
    agg = collections.OrderedDict()
    for n, t, g, b in sel:
        a = agg.setdefault(n, [0, 0.0, g, b]); a[0] += 1; a[1] += t
    print("### %s (%d launches, %.1f us of kernel time)" % (what, len(sel), tot))
    print("| kernel | launches | avg us | share | grid x block |\n|---|---|---|---|---|")
    for n, a in agg.items():
        print("| `%s` | %d | %.1f | %.1f %% | %s x %s |" % (n[:70], a[0], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
    print()
    return agg, tot


def is_x(n, mode):
    return re.match(r"XFusedH?<Fac<\d+, \d+>, %d" % mode, n) is not None


fwd = [r for r in rows if is_x(r[0], 2)]
adj = [r for r in rows if is_x(r[0], 3)]
# the loop kernels between the first and the last fused x launch of each kind
def span(mode):
    idx = [i for i, r in enumerate(rows) if is_x(r[0], mode)]
    return rows[max(idx[0] - 1, 0): idx[-1] + 3] if idx else []
fa, ft = table(span(2), "forward time loop, steady state (4 launches per step: y-inverse -> snapshot slot, fused x, y-forward, fused z step)") if fwd else ({}, 0)
aa, at = table(span(3), "adjoint time loop, steady state (4 launches per step; the forward state is read from its snapshot slot by the x pass)") if adj else ({}, 0)
if fwd and adj:
    nf, na = len(fwd), len(adj)
    xa = sum(r[1] for r in adj) / na
    step = ft / nf + at / na
    print("ncu: one forward step %.1f us + one adjoint step %.1f us = %.1f us; x-adj %.1f us per launch, share of the step pair %.3f." % (ft / nf, at / na, step, xa, xa / step))
