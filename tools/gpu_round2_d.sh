#!/bin/bash
# round-2 GPU session D (1 GPU): grid accumulation of the adjoint x pass (SMO_OPT_GRID_ACC) A/B, ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
SMO_KDYN_OPTS="7=1" python -m pytest tests/test_gpu_parity.py tests/test_gpu_history.py -m gpu -q -k "kdyn or config2 or reproducible" > $O/r2d_tests_gacc.log 2>&1; echo "tests(gacc) rc=$?" | tee -a $O/r2d_tests_gacc.log
for g in 0 1; do
  SMO_OPT_7=$g GRAPH=1 python tools/time_kdyn.py 128 50 > $O/r2d_time128_gacc$g.log 2>&1
  SMO_OPT_7=$g GRAPH=1 python tools/time_kdyn.py 256 10 > $O/r2d_time256_gacc$g.log 2>&1
done
SMO_KDYN_OPTS="7=1" timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:6XFusedINS_3FacILi16ELi12EEELi3 -c 2 -o $O/prof_r2d_xadj128_gacc -f python tools/ncu_kdyn.py 128 2 > $O/r2d_ncu128.log 2>&1; echo "ncu128 rc=$?"
SMO_KDYN_OPTS="7=1" timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:7XFusedHINS_3FacILi16ELi12EEELi3 -c 1 -o $O/prof_r2d_xadj256_gacc -f python tools/ncu_kdyn.py 256 1 > $O/r2d_ncu256.log 2>&1; echo "ncu256 rc=$?"
tail -3 $O/r2d_tests_gacc.log; grep -h "N=\|x-adj\|x-fwd" $O/r2d_time*.log
