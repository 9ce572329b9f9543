#!/bin/bash
# session r2zb (2 GPUs): why do the pushing kernels take local time + transfer time?  peer-store flavours / per-tile system fences
# (historical record: the -DSMO_PEER_ST / -DSMO_PUSH_TILE_FENCE build switches these variants used were measured and removed again)
O=gpurun_out; mkdir -p $O
P=${P:-2}
L=$O/r2zb_push_variants_p$P.txt
: > $L
for v in stsys stwt tfence; do
  echo "== library variant $v" >> $L
  SMO_B200_LIB=build/variants/libsmo_$v.so VARIANTS="1,1,0,1" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/time_kdyn_mp.py 128 200 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "rc=${PIPESTATUS[0]}" >> $L
done
echo "== default library, option 99 (peer buffers -> local: kernels without NVLink traffic, WRONG RESULTS)" >> $L
SMO_OPT_99=1 VARIANTS="1,1,0,1" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/time_kdyn_mp.py 128 200 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L
cat $L
