"""smallest run that exercises every warp-synchronous kernel path (dynamo 128^3: 2 steps each way; SH23 N=256: 3 steps, 6
instances) - meant for `compute-sanitizer --tool racecheck|memcheck python tools/sanitize_small.py` (development tool)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from spheremanopt_b200 import kdyn, sh23

dom = kdyn.Domain(128)
M = dom.M
g = torch.Generator(device="cuda").manual_seed(0)
B = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
U = torch.randn(3 * M ** 3, dtype=torch.float64, device="cuda", generator=g)
st = kdyn.GEN_BUFFER(128, dom, 2, checkpoint_every=0)
X = [kdyn.DevVec(B), kdyn.DevVec(U)]
f = kdyn.FWD_Solve_IVP_Lin(X, dom, 10.0, 1e-3, 2, 2, st)
gr = kdyn.ADJ_Solve_IVP_Lin(X, dom, 10.0, 1e-3, 2, 2, st)
fi = kdyn.FWD_Solve_IVP_Lin(X, dom, 10.0, 1e-3, 2, 2, st, "Integrated")
gi = kdyn.ADJ_Solve_IVP_Lin(X, dom, 10.0, 1e-3, 2, 2, st, "Integrated")
sd = sh23.Domain(256)
Xs = torch.randn(6 * sd.M, dtype=torch.float64, device="cuda", generator=g) * 0.05
ss = sh23.GEN_BUFFER(sd, 3, batch=6)
J = sh23.forward_batch(Xs, sd, 0.1, 3, ss)
G = sh23.adjoint_batch(sd, 0.1, 3, ss)
# the ensemble variant of the SH23 kernels (more than 2*148 CTAs: 128-register build, 8 CTAs per SM)
nb = 1300
Xe = torch.randn(nb * sd.M, dtype=torch.float64, device="cuda", generator=g) * 0.05
se = sh23.GEN_BUFFER(sd, 3, batch=nb)
Je = sh23.forward_batch(Xe, sd, 0.1, 3, se)
Ge = sh23.adjoint_batch(sd, 0.1, 3, se)
torch.cuda.synchronize()
print("ok", f, fi, float(J.sum()), float(G.abs().sum()), float(Je.sum()), float(Ge.abs().sum()))
