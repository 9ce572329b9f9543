#!/bin/bash
# session r2zc (P GPUs): time-stamp trace of the in-kernel hand-shakes (library variant built with -DSMO_XS_TRACE)
O=gpurun_out; mkdir -p $O
P=${P:-2}
L=$O/r2zc_trace_p$P.txt
: > $L
for opts in "" "99"; do
  echo "== trace, option 99 (no NVLink traffic, wrong results) = '$opts'" >> $L
  if [ -n "$opts" ]; then export SMO_OPT_99=1; fi
  SMO_B200_LIB=build/variants/libsmo_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/trace_mp.py 128 100 2>&1 | grep -v "^\[\|Warning\|warn\|^\*\|OMP_NUM\|^$" >> $L; echo "rc=${PIPESTATUS[0]}" >> $L
done
cat $L
