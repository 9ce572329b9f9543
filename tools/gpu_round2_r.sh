#!/bin/bash
# round-2 GPU session R (1 GPU): fused z step at 256^3 with one line per warp (build/libsmo_zlp.so) vs the CTA-barrier version
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== default build (24 threads per line, CTA-wide barriers)" > $O/r2r_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|z-step" >> $O/r2r_time256.log
echo "== one line per warp (build/libsmo_zlp.so)" >> $O/r2r_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 build/libsmo_zlp.so 2>&1 | grep "N=\|z-step" >> $O/r2r_time256.log
SMO_B200_LIB=$PWD/build/libsmo_zlp.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_grids or reproducible or kdyn_f_gradf" > $O/r2r_tests_zlp.log 2>&1; echo "tests(zlp) rc=$?" | tee -a $O/r2r_tests_zlp.log
cat $O/r2r_time256.log; tail -3 $O/r2r_tests_zlp.log
