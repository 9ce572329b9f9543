#!/bin/bash
# session r2z3: PDL at the small grids (early trigger = default build, late trigger variant), one GPU
# (historical record: the -DSMO_PDL_LATE / -DSMO_STAGGER_NS build switches these variants used were measured and removed again)
O=gpurun_out; mkdir -p $O
echo "== default build (early trigger)" >> $O/r2z3_pdl_small.txt
timeout 300 python tools/ab_pdl.py 16:1000 24:1000 64:300 >> $O/r2z3_pdl_small.txt 2>&1
echo "== variant late" >> $O/r2z3_pdl_small.txt
SMO_LIB=build/variants/libsmo_late.so timeout 300 python tools/ab_pdl.py 24:1000 64:300 >> $O/r2z3_pdl_small.txt 2>&1
grep -v identical $O/r2z3_pdl_small.txt
