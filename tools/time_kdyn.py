"""ad-hoc timing of the dynamo time loops on one GPU, with the per-kernel-class split (development tool)"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from spheremanopt_b200 import _cabi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if len(sys.argv) > 3:
    _cabi.LIB_PATH = sys.argv[3]   # a tuning variant built by tools/variants.py
    print("library:", sys.argv[3])
from spheremanopt_b200 import kdyn
dom = kdyn.Domain(N)
lib = dom.lib
M = dom.M
g = torch.Generator(device="cuda").manual_seed(0)
B = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
U = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
# make them band limited and unit norm through the library itself
B = kdyn.to_grid(dom, kdyn.to_coef(dom, B)); U = kdyn.to_grid(dom, kdyn.to_coef(dom, U))
ip = lambda a: kdyn.Inner_Prod_3(kdyn.DevVec(a), kdyn.DevVec(a), dom)
B = B / np.sqrt(ip(B)); U = U / np.sqrt(ip(U))
st = kdyn.GEN_BUFFER(N, dom, nit)
X = [kdyn.DevVec(B), kdyn.DevVec(U)]
args = (dom, 10.0, 1e-3, nit, nit, st)
names = {0: "total", 1: "z-pass", 2: "y-pass", 3: "x-fwd", 4: "epilogue", 5: "a2a", 6: "x-adj", 7: "z-step"}
C_ = (N // 2) * (N - 1) ** 2 * 16; P1 = (N // 2) * (N - 1) * M * 16; P2 = (N // 2) * M * M * 16
alg_f = 9 * C_ + 12 * P1 + 15 * P2; alg_a = 18 * C_ + 24 * P1 + 27 * P2
import os
chunk_sets = [tuple(int(v) for v in cs.split(",")) for cs in os.environ.get("CHUNKS", "1,1").split(";")]
for k, v in os.environ.items():
    if k.startswith("SMO_OPT_"):
        lib.smo_kdyn_set_option(dom.h, int(k[8:]), int(v))
if os.environ.get("GRAPH"):
    lib.smo_kdyn_use_graph(dom.h, 1)
for _ in range(3):
    kdyn.FWD_Solve_IVP_Lin(X, *args); kdyn.ADJ_Solve_IVP_Lin(X, *args)   # warm-up (lazy module loading, graph capture)
for cf, ca in chunk_sets:
  lib.smo_kdyn_set_chunks(dom.h, cf, ca)
  print("chunks fwd/adj:", cf, ca)
  for which in ((0, 1, 2, 3, 6, 4, 7) if len(chunk_sets) == 1 and not os.environ.get("TOTAL_ONLY") else (0,) * int(os.environ.get("TOTAL_ONLY") or 1)):
    for fn, nm, alg in ((kdyn.FWD_Solve_IVP_Lin, "fwd", alg_f), (kdyn.ADJ_Solve_IVP_Lin, "adj", alg_a)):
        lib.smo_kdyn_profile_set(dom.h, which)
        torch.cuda.synchronize(); t = time.time()
        fn(X, *args)
        torch.cuda.synchronize(); dt = time.time() - t
        ms = C.c_double(); n = C.c_longlong()
        lib.smo_kdyn_profile_read(dom.h, C.byref(ms), C.byref(n))
        if which == 0:
            print("%s N=%d: %.3f ms/step  (%.1f GB/s algorithmic, %.1f%% of 6446)" % (nm, N, dt / nit * 1e3, alg * nit / dt / 1e9, alg * nit / dt / 1e9 / 64.463))
        elif n.value:
            print("   %s %-8s %8.3f ms/step over %d launches (%.3f ms/launch)" % (nm, names[which], ms.value / nit, n.value, ms.value / max(n.value, 1)))
