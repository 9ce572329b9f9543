#!/bin/bash
# round-2 GPU session K (1 GPU): TMA tensor copies for the spectral tiles (SMO_OPT_TMA_SIN) - experimental build under build/
cd "$(dirname "$0")/.."
O=gpurun_out
export SMO_B200_LIB=$PWD/build/libsmo_exp.so
SMO_KDYN_OPTS="9=1" timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kdyn_f_gradf or large_grids or checkpointed or integrated or graph" > $O/r2k_tests_tma.log 2>&1; echo "tests(9=1) rc=$?" | tee -a $O/r2k_tests_tma.log
for o in 0 1; do
  echo "== tma_sin=$o" >> $O/r2k_time.log
  SMO_OPT_9=$o GRAPH=1 timeout 300 python tools/time_kdyn.py 128 50 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2k_time.log
  SMO_OPT_9=$o GRAPH=1 timeout 300 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2k_time.log
done
tail -5 $O/r2k_tests_tma.log; cat $O/r2k_time.log
