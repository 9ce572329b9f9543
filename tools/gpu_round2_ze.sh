#!/bin/bash
# session r2ze (P GPUs): final hand-shake (device-scope CTA fences) - parity vs the oracle and the time-stamp trace at P ranks
O=gpurun_out; mkdir -p $O
P=${P:-8}
L=$O/r2ze_final_handshake_p$P.txt
: > $L
echo "== parity (default library), P=$P" >> $L
MP_CASES="${MP_CASES:-32:6,128:2}" MP_VARIANTS="3" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "parity rc=${PIPESTATUS[0]}" >> $L
echo "== trace (library built with -DSMO_XS_TRACE, otherwise the default), 128^3 x 200 steps" >> $L
SMO_B200_LIB=build/variants/libsmo_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/trace_mp.py 128 200 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "rc=${PIPESTATUS[0]}" >> $L
cat $L
