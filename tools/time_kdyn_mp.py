"""multi-GPU timing of the dynamo time loops with the per-kernel-class split (development tool)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 tools/time_kdyn_mp.py [N] [nit]

VARIANTS="waves,chunks,two_streams[,split[,bulk_push]];..." runs several tuning variants in ONE process (default "1,1,0"); split=1 adds
the per-kernel-class times.  J is printed for every variant (must agree between variants and GPU counts).
"""
import ctypes as C
import os
import sys
import time

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
M = dom.M
g = torch.Generator(device="cuda").manual_seed(rank)
B = torch.randn(3 * M * M * dom.nz, dtype=torch.float64, device="cuda", generator=g)
U = torch.randn(3 * M * M * dom.nz, dtype=torch.float64, device="cuda", generator=g)
B = kdyn.to_grid(dom, kdyn.to_coef(dom, B)); U = kdyn.to_grid(dom, kdyn.to_coef(dom, U))
ip = lambda a: kdyn.Inner_Prod_3(kdyn.DevVec(a), kdyn.DevVec(a), dom)
B = B / np.sqrt(ip(B)); U = U / np.sqrt(ip(U))
st = kdyn.GEN_BUFFER(N, dom, nit, checkpoint_every=int(os.environ.get("CKPT", "0")))
X = [kdyn.DevVec(B), kdyn.DevVec(U)]
args = (dom, 10.0, 1e-3, nit, nit, st)
names = {0: "total", 1: "z-pass", 2: "y-pass", 3: "x-fwd", 5: "a2a", 6: "x-adj", 7: "z-step"}
for k, v in os.environ.items():
    if k.startswith("SMO_OPT_"):
        lib.smo_kdyn_set_option(dom.h, int(k[8:]), int(v))
if os.environ.get("GRAPH", "1") != "0":
    lib.smo_kdyn_use_graph(dom.h, 1)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for spec in os.environ.get("VARIANTS", "1,1,0,1").split(";"):
    v = [int(x) for x in spec.split(",")]
    waves, chunks, two = v[0], v[1], v[2]
    split = v[3] if len(v) > 3 else 0
    bulk = v[4] if len(v) > 4 else None
    if bulk is not None:
        lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_BULK_PUSH, bulk)
    lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PUSH_WAVES, waves)
    lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_TWO_STREAMS, two)
    lib.smo_kdyn_set_chunks(dom.h, chunks, chunks)
    lib.smo_kdyn_profile_set(dom.h, 0)
    for _ in range(3):
        f = kdyn.FWD_Solve_IVP_Lin(X, *args); kdyn.ADJ_Solve_IVP_Lin(X, *args)
    if rank == 0:
        print("== P=%d N=%d waves=%d chunks=%d two_streams=%d bulk_push=%s: J=%.15e" % (world, N, waves, chunks, two, bulk, -f), flush=True)
    for which in ((0, 0, 2, 3, 6, 7, 5) if split else (0, 0)):
        for fn, nm in ((kdyn.FWD_Solve_IVP_Lin, "fwd"), (kdyn.ADJ_Solve_IVP_Lin, "adj")):
            lib.smo_kdyn_profile_set(dom.h, which)
            if which:
                fn(X, *args)     # (a new profile kind is a new graph key: eager + capture first)
                fn(X, *args)
                lib.smo_kdyn_profile_set(dom.h, which)
            sync(); t = time.time()
            fn(X, *args)
            sync(); dt = time.time() - t
            ms = C.c_double(); n = C.c_longlong()
            lib.smo_kdyn_profile_read(dom.h, C.byref(ms), C.byref(n))
            if rank == 0:
                if which == 0:
                    print("   %s: %.1f us/step" % (nm, dt / nit * 1e6), flush=True)
                elif n.value:
                    print("      %s %-8s %8.1f us/step over %d launches (%.1f us/launch)" % (nm, names[which], ms.value / nit * 1e3, n.value, ms.value / max(n.value, 1) * 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
