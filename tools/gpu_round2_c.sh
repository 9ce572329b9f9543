#!/bin/bash
# round-2 GPU session C (2 GPUs): multi-GPU parity of every transport / chunking variant, then timing of the overlap variants
cd "$(dirname "$0")/.."
O=gpurun_out
P=${P:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py > $O/r2c_mp_parity_p$P.log 2>&1; echo "mp_parity rc=$?"
VARIANTS="${VARIANTS:-1,1,0,1;2,1,0;3,1,0;1,2,0;1,2,1,1;2,2,1;1,3,1;2,3,1;2,4,1}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/time_kdyn_mp.py 128 200 > $O/r2c_time_mp_p$P.log 2>&1; echo "time_mp rc=$?"
grep -v "^\[" $O/r2c_mp_parity_p$P.log | tail -40; cat $O/r2c_time_mp_p$P.log | grep -v Warning | tail -80
