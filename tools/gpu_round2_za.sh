#!/bin/bash
# session r2za (2 GPUs): peer pushes through TMA bulk stores of staged blocks (SMO_OPT_BULK_PUSH): parity, then A/B timing
O=gpurun_out; mkdir -p $O
P=${P:-2}
L=$O/r2za_bulk_push_p$P.txt
echo "# parity, SMO_KDYN_OPTS=11=3 (staged z-step + y-forward pushes), 128^3 x 2 steps, default transport + checkpointed" > $L
SMO_KDYN_OPTS="11=3" MP_CASES="128:2" MP_VARIANTS="3,10" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py 2>&1 | grep -v "^\[\|Warning\|warn" >> $L; echo "parity128 rc=${PIPESTATUS[0]}" >> $L
echo "# parity, SMO_KDYN_OPTS=11=3, standard small cases, push variants" >> $L
SMO_KDYN_OPTS="11=3" MP_VARIANTS="3,2,6,8,10" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29535 tests/mp_parity.py 2>&1 | grep -v "^\[\|Warning\|warn" >> $L; echo "parity small rc=${PIPESTATUS[0]}" >> $L
echo "# timing 128^3 x 200 steps: bulk_push 0 / 1 (z step) / 2 (y forward) / 3 (both)" >> $L
VARIANTS="1,1,0,1,0;1,1,0,1,1;1,1,0,1,2;1,1,0,1,3;1,1,0,0,0" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/time_kdyn_mp.py 128 200 2>&1 | grep -v "^\[\|Warning\|warn" >> $L; echo "timing rc=${PIPESTATUS[0]}" >> $L
cat $L
