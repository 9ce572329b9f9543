#!/bin/bash
# session r2zd (P GPUs): per-CTA device-scope fence + one system fence in the last CTA (variant) - parity and hand-shake trace
# (historical record: the variant of this session - device-scope fence per CTA - became the default; -DSMO_XSYNC_CTA_SYS restores the old protocol)
O=gpurun_out; mkdir -p $O
P=${P:-2}
L=$O/r2zd_gpufence_p$P.txt
: > $L
echo "== parity (variant library), 128^3 x 2 steps + small cases" >> $L
SMO_B200_LIB=build/variants/libsmo_trace_gpufence.so MP_CASES="128:2,32:6,64:3" MP_VARIANTS="3,0,10" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "parity rc=${PIPESTATUS[0]}" >> $L
echo "== trace (variant library)" >> $L
SMO_B200_LIB=build/variants/libsmo_trace_gpufence.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/trace_mp.py 128 100 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "rc=${PIPESTATUS[0]}" >> $L
cat $L
