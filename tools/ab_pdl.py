"""A/B timing of SMO_OPT_PDL (programmatic dependent launches inside the time loops) on one GPU, one process, alternating
settings on the same handle (development tool).  usage: python tools/ab_pdl.py [N:nit ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import os
from spheremanopt_b200 import _cabi
if os.environ.get("SMO_LIB"):
    _cabi.LIB_PATH = os.environ["SMO_LIB"]   # a tuning variant built by tools/variants.py
    print("library:", _cabi.LIB_PATH)
from spheremanopt_b200 import kdyn

cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(128, 200), (256, 12), (24, 1000)]
for N, nit in cases:
    dom = kdyn.Domain(N)
    lib, M = dom.lib, dom.M
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
    U = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
    B = kdyn.to_grid(dom, kdyn.to_coef(dom, B)); U = kdyn.to_grid(dom, kdyn.to_coef(dom, U))
    ip = lambda a: kdyn.Inner_Prod_3(kdyn.DevVec(a), kdyn.DevVec(a), dom)
    B = B / np.sqrt(ip(B)); U = U / np.sqrt(ip(U))
    st = kdyn.GEN_BUFFER(N, dom, nit, checkpoint_every=0)
    X = [kdyn.DevVec(B), kdyn.DevVec(U)]
    args = (dom, 10.0, 1e-3, nit, nit, st)
    lib.smo_kdyn_use_graph(dom.h, 1)
    ref = None
    for pdl in (0, 1):
        lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PDL, pdl)
        for _ in range(3):       # eager, capture, replay
            f = kdyn.FWD_Solve_IVP_Lin(X, *args); gr = kdyn.ADJ_Solve_IVP_Lin(X, *args)
        dig = (f, float(ip(gr[0].t)), float(ip(gr[1].t)))
        if ref is None:
            ref = dig
        print("N=%d pdl=%d  f=%.17g |gB|^2=%.17g |gU|^2=%.17g  identical=%s" % ((N, pdl) + dig + (dig == ref,)), flush=True)
    times = {0: [], 1: []}
    for rep in range(4):
        for pdl in (0, 1):
            lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PDL, pdl)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            torch.cuda.synchronize()
            e[0].record(); kdyn.FWD_Solve_IVP_Lin(X, *args)
            e[1].record(); kdyn.ADJ_Solve_IVP_Lin(X, *args)
            e[2].record(); torch.cuda.synchronize()
            times[pdl].append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    for pdl in (0, 1):
        fw = sorted(t[0] for t in times[pdl]); ad = sorted(t[1] for t in times[pdl])
        print("N=%d nit=%d pdl=%d : fwd %.3f ms (min) %.3f (max) = %.2f us/step | adj %.3f ms (min) %.3f (max) = %.2f us/step | pair %.3f ms"
              % (N, nit, pdl, fw[0], fw[-1], fw[0] / nit * 1e3, ad[0], ad[-1], ad[0] / nit * 1e3, fw[0] + ad[0]), flush=True)
    del st, X, B, U, dom
    torch.cuda.empty_cache()
