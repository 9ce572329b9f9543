#!/bin/bash
# round-2 GPU session Q (1 GPU): 256^3 forward step - fused z step with 1 line per CTA (build/libsmo_tzs1.so) vs 2; ncu of the half-length forward x pass
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== default (2 lines per CTA)" > $O/r2q_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" >> $O/r2q_time256.log
echo "== 1 line per CTA (build/libsmo_tzs1.so)" >> $O/r2q_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 build/libsmo_tzs1.so 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" >> $O/r2q_time256.log
SMO_B200_LIB=$PWD/build/libsmo_tzs1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_grids" > $O/r2q_tests_tzs1.log 2>&1; echo "tests(tzs1) rc=$?" | tee -a $O/r2q_tests_tzs1.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:7XFusedHINS_3FacILi16ELi12EEELi2 -c 1 -o $O/prof_r2q_xfwd256 -f python tools/ncu_kdyn.py 256 1 > $O/r2q_ncu256.log 2>&1; echo "ncu256 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:5ZStepINS_3FacILi24 -c 1 -o $O/prof_r2q_zstep256 -f python tools/ncu_kdyn.py 256 1 > $O/r2q_ncu256z.log 2>&1; echo "ncu256z rc=$?"
cat $O/r2q_time256.log; tail -3 $O/r2q_tests_tzs1.log
