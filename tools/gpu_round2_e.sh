#!/bin/bash
# round-2 GPU session E (1 GPU): A/B of grid accumulation (phase-5 read-modify-write) and the TMA bulk copy of the velocity tile
cd "$(dirname "$0")/.."
O=gpurun_out
SMO_KDYN_OPTS="7=1,8=1" python -m pytest tests/test_gpu_parity.py tests/test_gpu_history.py -m gpu -q -k "kdyn or config2 or reproducible" > $O/r2e_tests_opts.log 2>&1; echo "tests(7=1,8=1) rc=$?" | tee -a $O/r2e_tests_opts.log
for o in "0 0" "1 0" "0 1" "1 1"; do
  set -- $o
  echo "== grid_acc=$1 bulk_u=$2" >> $O/r2e_time.log
  SMO_OPT_7=$1 SMO_OPT_8=$2 GRAPH=1 python tools/time_kdyn.py 128 50 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2e_time.log
  SMO_OPT_7=$1 SMO_OPT_8=$2 GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd" >> $O/r2e_time.log
done
tail -3 $O/r2e_tests_opts.log; cat $O/r2e_time.log
