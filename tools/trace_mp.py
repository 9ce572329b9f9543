"""development tool: time-stamp trace of the in-kernel hand-shakes of the multi-GPU time loops (library built with -DSMO_XS_TRACE)

    SMO_B200_LIB=build/variants/libsmo_trace.so python -m torch.distributed.run --nnodes=1 --nproc-per-node P \
        --master-addr 127.0.0.1 --master-port 29511 tools/trace_mp.py [N] [nit]

Per traced launch (kernels that wait for / signal peers: inverse y pass, forward y pass, fused z step) CTA 0 records globaltimer at
kernel entry, after the wait, after its work loop and after its fence + count; the last CTA to finish records when it arrived and
when it had published the flags.  Prints the average stage lengths per kernel class over the middle of one forward and one adjoint
solve (every rank analyses its own GPU clock; rank 0 and the last rank print)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from spheremanopt_b200 import _cabi, kdyn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dom = kdyn.Domain(N, device="cuda:%d" % local)
lib = dom.lib
raw = C.CDLL(_cabi.LIB_PATH)
raw.smo_debug_xs_trace.argtypes = [C.c_void_p, C.c_int]
M = dom.M
g = torch.Generator(device="cuda").manual_seed(rank)
B = torch.randn(3 * M * M * dom.nz, dtype=torch.float64, device="cuda", generator=g)
U = torch.randn(3 * M * M * dom.nz, dtype=torch.float64, device="cuda", generator=g)
B = kdyn.to_grid(dom, kdyn.to_coef(dom, B)); U = kdyn.to_grid(dom, kdyn.to_coef(dom, U))
ip = lambda a: kdyn.Inner_Prod_3(kdyn.DevVec(a), kdyn.DevVec(a), dom)
B = B / np.sqrt(ip(B)); U = U / np.sqrt(ip(U))
st = kdyn.GEN_BUFFER(N, dom, nit, checkpoint_every=0)
X = [kdyn.DevVec(B), kdyn.DevVec(U)]
args = (dom, 10.0, 1e-3, nit, nit, st)
for k, v in os.environ.items():
    if k.startswith("SMO_OPT_"):
        lib.smo_kdyn_set_option(dom.h, int(k[8:]), int(v))
if os.environ.get("GRAPH", "1") != "0":
    lib.smo_kdyn_use_graph(dom.h, 1)
for _ in range(3):
    f = kdyn.FWD_Solve_IVP_Lin(X, *args); gr = kdyn.ADJ_Solve_IVP_Lin(X, *args)
gn = [float(ip(v.t)) for v in gr]
if rank == 0:     # (same inputs as tools/time_kdyn_mp.py: J must agree with its runs at the same GPU count)
    print("P=%d N=%d nit=%d: J=%.15e  <gB,gB>=%.15e  <gU,gU>=%.15e" % (world, N, nit, -f, gn[0], gn[1]), flush=True)
NS = 8192
out = []
names = {128004: "y pass", 128006: "y-fwd (staged)", 96009: "z step"}
for fn, nm in ((kdyn.FWD_Solve_IVP_Lin, "forward"), (kdyn.ADJ_Solve_IVP_Lin, "adjoint")):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    raw.smo_debug_xs_trace(None, 1)
    torch.cuda.synchronize()
    fn(X, *args)
    torch.cuda.synchronize()
    buf = np.zeros(NS * 8, dtype=np.uint64)
    raw.smo_debug_xs_trace(buf.ctypes.data, 0)
    t = buf.reshape(NS, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t = t[np.argsort(t[:, 0])]
    if rank not in (0, world - 1):
        continue
    n = len(t)
    lo, hi = n // 4, 3 * n // 4      # the middle of the time loop
    out.append("rank %d %s solve: %d traced launches, loop period %.2f us per step" % (rank, nm, n, (t[hi, 0] - t[lo, 0]) / 1e3 / ((hi - lo) / 3.0)))
    cls = {}
    for i in range(lo, hi):
        r = t[i]
        waits = r[1] - r[0]; work = r[2] - r[1]
        sig = r[5] > 0
        key = (int(r[7]), bool(sig), "waits" if waits > 2500 else "-")
        end = r[5] if sig else r[2]
        d = cls.setdefault(key, [])
        d.append((waits, work, (r[3] - r[2]) if r[3] > 0 else 0, (r[4] - r[2]) if sig else 0, (r[5] - r[4]) if sig else 0, end - r[0], t[i + 1, 0] - end, t[i + 1, 0] - r[0]))
    out.append("   %-28s %6s | %8s %8s %8s %10s %8s | %8s %8s %8s" % ("kernel (signals?, waited?)", "n", "entry>go", "cta0work", "cta0fnc", "last-cta0", "lastflag", "total", "gap>next", "period"))
    for key, d in sorted(cls.items()):
        a = np.array(d, dtype=np.float64).mean(axis=0) / 1e3
        out.append("   %-28s %6d | %8.2f %8.2f %8.2f %10.2f %8.2f | %8.2f %8.2f %8.2f" % ("%s %s %s" % (names.get(key[0], key[0]), "signals" if key[1] else "", key[2]), len(d), *a))
for r in sorted({0, world - 1}):
    if rank == r:
        print("\n".join(out), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
