#!/bin/bash
# session r2z: programmatic dependent launch (SMO_OPT_PDL) - bitwise test + A/B timing at 128^3 / 256^3 / 24^3, one GPU
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r2z_gpu.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "programmatic or graph_replay" > $O/r2z_tests_pdl.log 2>&1; echo "tests rc=$?" | tee -a $O/r2z_tests_pdl.log
timeout 500 python tools/ab_pdl.py 128:200 256:12 24:1000 > $O/r2z_pdl_ab.txt 2>&1; echo "ab rc=$?" | tee -a $O/r2z_pdl_ab.txt
tail -5 $O/r2z_tests_pdl.log; cat $O/r2z_pdl_ab.txt
