#!/bin/bash
# round-2 GPU session F (1 GPU): full GPU suite with the new defaults, 256^3 timing after the whole-line tile fix, ensemble of optimisations
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/r2f_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2f_tests.log
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" > $O/r2f_time256.log
GRAPH=1 python tools/time_kdyn.py 128 50 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" > $O/r2f_time128.log
SMO_ENS_TOTAL=512 SMO_ENS_ITERS=20 timeout 900 python bench.py --workload sh23opt --steps 1 --warmup 0 > $O/r2f_bench_sh23opt512.json 2> $O/r2f_bench_sh23opt512.err; echo "sh23opt rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2f_bench_kdyn128.json 2> $O/r2f_bench_kdyn128.err; echo "bench128 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:6XFusedINS_3FacILi16ELi12EEELi3 -c 2 -o $O/prof_r2f_xadj128 -f python tools/ncu_kdyn.py 128 2 > $O/r2f_ncu128.log 2>&1; echo "ncu128 rc=$?"
tail -3 $O/r2f_tests.log; cat $O/r2f_time256.log $O/r2f_time128.log; tail -c 1500 $O/r2f_bench_sh23opt512.json; tail -5 $O/r2f_bench_sh23opt512.err
