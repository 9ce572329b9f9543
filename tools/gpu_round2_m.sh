#!/bin/bash
# round-2 GPU session M (1 GPU): half-length x pass with T = 2 columns per CTA (build/libsmo_xh2.so) vs T = 4; SH23 adjoint run-to-run spread
cd "$(dirname "$0")/.."
O=gpurun_out
for i in 1 2 3 4; do python tools/time_sh23ens.py 4096 >> $O/r2m_sh23_spread.log 2>&1; done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv >> $O/r2m_sh23_spread.log
echo "== T=4 (default build)" > $O/r2m_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2m_time256.log
echo "== T=2 (build/libsmo_xh2.so)" >> $O/r2m_time256.log
GRAPH=1 python tools/time_kdyn.py 256 10 build/libsmo_xh2.so 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2m_time256.log
SMO_B200_LIB=$PWD/build/libsmo_xh2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_grids" > $O/r2m_tests_xh2.log 2>&1; echo "tests(xh2) rc=$?" | tee -a $O/r2m_tests_xh2.log
cat $O/r2m_sh23_spread.log $O/r2m_time256.log; tail -3 $O/r2m_tests_xh2.log
