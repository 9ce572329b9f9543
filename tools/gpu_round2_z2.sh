#!/bin/bash
# session r2z2: PDL trigger placement (early / late) and explicit start stagger of the y / z kernels, one GPU
# (historical record: the -DSMO_PDL_LATE / -DSMO_STAGGER_NS build switches these variants used were measured and removed again)
O=gpurun_out; mkdir -p $O
for v in late stag500 stag1000; do
  echo "== variant $v" >> $O/r2z2_pdl_variants.txt
  SMO_LIB=build/variants/libsmo_$v.so timeout 300 python tools/ab_pdl.py 128:200 48:400 >> $O/r2z2_pdl_variants.txt 2>&1; echo "rc=$?" >> $O/r2z2_pdl_variants.txt
done
echo "== default build" >> $O/r2z2_pdl_variants.txt
timeout 300 python tools/ab_pdl.py 32:600 48:400 64:300 >> $O/r2z2_pdl_variants.txt 2>&1
cat $O/r2z2_pdl_variants.txt
