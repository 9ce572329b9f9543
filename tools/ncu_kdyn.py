"""short dynamo run for profiling under ncu: one forward + one adjoint of `nit` steps at Npts^3 (development tool)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from spheremanopt_b200 import kdyn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dom = kdyn.Domain(N)
M = dom.M
g = torch.Generator(device="cuda").manual_seed(0)
B = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
U = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
st = kdyn.GEN_BUFFER(N, dom, nit)
X = [kdyn.DevVec(B), kdyn.DevVec(U)]
f = kdyn.FWD_Solve_IVP_Lin(X, dom, 10.0, 1e-3, nit, nit, st)
gr = kdyn.ADJ_Solve_IVP_Lin(X, dom, 10.0, 1e-3, nit, nit, st)
torch.cuda.synchronize()
print("f =", f, " |gradB|^2 =", kdyn.Inner_Prod_3(gr[0], gr[0], dom))
