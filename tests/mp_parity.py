"""Multi-GPU parity check of the slab-decomposed dynamo path against the oracle (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py

Every transport is exercised: peer-memory transposes with in-kernel hand-shakes (default), peer-memory transposes with
barrier launches, grouped ncclSend/ncclRecv; CUDA-graph replay; host vectors (Mode H) and DevVec (Mode D).
Rank 0 prints "MP_PARITY OK" on success.  tests/test_gpu_multi.py wraps this for pytest -m gpu on boxes with >= 2 GPUs.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kdyn as okd                      # noqa: E402  (the checker)
from spheremanopt_b200 import _cabi, kdyn          # noqa: E402
from spheremanopt_b200.devvec import DevVec        # noqa: E402
from tests.common import kdyn_field, relerr        # noqa: E402

TOL = 1e-9


def main():
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [(32, 6)] if world > 2 else [(16, 5), (32, 6), (64, 3)]      # (64: nz = 48 per rank at P = 2 -> chunked variants run)
    if os.environ.get("MP_CASES"):      # e.g. "128:2" - the full grid of config 3 (the staged z-step push exists for 128^3 / 256^3 only)
        cases = [tuple(int(v) for v in c.split(":")) for c in os.environ["MP_CASES"].split(",")]
    worst = 0.0
    for Npts, nit in cases:
        od = okd.domain_kdyn(Npts)
        if (Npts // 2) % world or (3 * Npts // 2) % world:
            continue
        B0, U = kdyn_field(od, 1), kdyn_field(od, 2)
        Rm, dt = 1.0, 1e-3
        D = okd.GEN_BUFFER(Npts, od, nit)
        fo = okd.FWD_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
        go = okd.ADJ_Solve_IVP_Lin([B0, U], od, Rm, dt, nit, nit, D)
        # (peer memory, in-kernel hand-shakes, graph replay, pull transposes, z chunks of the y-x-y section, two streams,
        #  work items per CTA of the pushing kernels, checkpoint spacing)
        variants = ((True, 1, 1, 1, 1, 0, 1, 0), (True, 0, 0, 1, 1, 0, 1, 0), (True, 1, 0, 0, 1, 0, 1, 0), (True, 1, 1, 0, 1, 0, 1, 0),
                    (True, 0, 0, 0, 1, 0, 1, 0), (False, 0, 0, 0, 1, 0, 1, 0),
                    (True, 1, 1, 0, 2, 1, 2, 0), (True, 1, 0, 0, 2, 1, 1, 0), (True, 1, 1, 0, 2, 0, 3, 0), (True, 1, 1, 1, 2, 1, 1, 0),
                    (True, 1, 1, 0, 1, 0, 1, 4), (True, 1, 1, 0, 2, 1, 2, 3), (False, 0, 0, 0, 1, 0, 1, 4))
        if os.environ.get("MP_VARIANTS"):   # subset by index, e.g. "3,10"
            variants = tuple(variants[int(i)] for i in os.environ["MP_VARIANTS"].split(","))
        for peer, ksync, fused, pull, chunks, two, waves, every in variants:
            if chunks > 1 and ((3 * Npts // 2) // world) % (8 * chunks):
                continue     # (a z chunk must hold whole tiles of the y passes)
            dom = kdyn.Domain(Npts, device="cuda:%d" % local, peer_memory=peer)
            dom.lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_KERNEL_SYNC, ksync)
            dom.lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PEER_PULL, pull)
            dom.lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_TWO_STREAMS, two)
            dom.lib.smo_kdyn_set_option(dom.h, _cabi.SMO_OPT_PUSH_WAVES, waves)
            dom.lib.smo_kdyn_set_chunks(dom.h, chunks, chunks)
            store = kdyn.GEN_BUFFER(Npts, dom, nit, checkpoint_every=every)
            if peer and ksync and fused:
                dom.lib.smo_kdyn_use_graph(dom.h, 1)     # graph replay with device-side epoch bases (3rd call onwards)
            for rep in range(4):     # repeatedly: epochs / counters / graph replays must survive repeated calls
                f = kdyn.FWD_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
                g = kdyn.ADJ_Solve_IVP_Lin([B0, U], dom, Rm, dt, nit, nit, store)
            e = [abs(f - fo) / abs(fo), relerr(g[0], go[0]), relerr(g[1], go[1])]
            # device-resident vectors (sharded DevVec)
            Xd = [DevVec(dom.slab_from_host(B0)), DevVec(dom.slab_from_host(U))]
            fd = kdyn.FWD_Solve_IVP_Lin(Xd, dom, Rm, dt, nit, nit, store)
            gd = kdyn.ADJ_Solve_IVP_Lin(Xd, dom, Rm, dt, nit, nit, store)
            e += [abs(fd - fo) / abs(fo), relerr(dom.host_from_slab(gd[0].t), go[0]), relerr(dom.host_from_slab(gd[1].t), go[1])]
            ip = kdyn.Inner_Prod_3(gd[0], Xd[0], dom)
            e.append(abs(ip - okd.Inner_Prod_3(go[0], B0, od)) / abs(okd.Inner_Prod_3(go[0], B0, od)))
            worst = max(worst, max(e))
            if rank == 0:
                print("P=%d N=%d peer=%d ksync=%d graph=%d pull=%d chunks=%d two_streams=%d waves=%d ckpt_every=%d: max rel.err %.2e"
                      % (world, Npts, peer, ksync, fused, pull, chunks, two, waves, every, max(e)), flush=True)
            del dom, store
    ok = torch.tensor([1 if worst <= TOL else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MP_PARITY OK" if int(ok.item()) == 1 else "MP_PARITY FAILED (worst %.3e)" % worst, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
