"""ad-hoc timing of the batched SH23 solves (development tool): python tools/time_sh23ens.py [batch] [lib.so]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spheremanopt_b200 import _cabi
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if len(sys.argv) > 2:
    _cabi.LIB_PATH = sys.argv[2]
    if "r1" in sys.argv[2]:      # older ABI: bind only what exists
        for k in ("smo_vec_checksum", "smo_vec_dot_dev", "smo_microbench_dfma"):
            _cabi.SIGNATURES.pop(k, None)
from spheremanopt_b200 import sh23
dom, X0 = sh23.Generate_IC(0.0725)
M0 = np.linspace(0.05, 0.1, batch)
X = torch.from_numpy(np.sqrt(M0 / 0.0725)[:, None] * X0[None, :]).to(dom.device).reshape(-1).contiguous()
store = sh23.GEN_BUFFER(dom, 500, 256, batch=batch)
for _ in range(3):
    sh23.forward_batch(X, dom, 0.1, 500, store); sh23.adjoint_batch(dom, 0.1, 500, store)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf, ta = [], []
for _ in range(5):
    ev[0].record(); J = sh23.forward_batch(X, dom, 0.1, 500, store)
    ev[1].record(); G = sh23.adjoint_batch(dom, 0.1, 500, store)
    ev[2].record(); torch.cuda.synchronize()
    tf.append(ev[0].elapsed_time(ev[1])); ta.append(ev[1].elapsed_time(ev[2]))
print("lib=%s batch=%d: forward %.3f ms  adjoint %.3f ms  (min %.3f / %.3f)  J0=%.12e |G|=%.12e" % (_cabi.LIB_PATH.split("/")[-1], batch, np.mean(tf), np.mean(ta), min(tf), min(ta), float(J[0]), float(G.norm())))
