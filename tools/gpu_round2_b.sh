#!/bin/bash
# round-2 GPU session B (1 GPU): half-length x pass at 256^3, SH23 A/B against the round-1 library, bench lines, ncu of the 256^3 adjoint step
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2b_tests.log
python tools/time_sh23ens.py 4096 > $O/r2b_sh23ens_ab.log 2>&1
python tools/time_sh23ens.py 4096 build/libsmo_r1.so >> $O/r2b_sh23ens_ab.log 2>&1
python tools/time_sh23ens.py 2368 >> $O/r2b_sh23ens_ab.log 2>&1
GRAPH=1 python tools/time_kdyn.py 256 10 > $O/r2b_time256.log 2>&1; echo "time256 rc=$?"
python bench.py --workload sh23 --steps 5 --warmup 3 > $O/r2b_bench_sh23.json 2> $O/r2b_bench_sh23.err; echo "sh23 rc=$?"
python bench.py --workload kdyn24 --steps 5 --warmup 3 > $O/r2b_bench_kdyn24.json 2> $O/r2b_bench_kdyn24.err; echo "kdyn24 rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2b_bench_kdyn128.json 2> $O/r2b_bench_kdyn128.err; echo "bench128 rc=$?"
timeout 600 python bench.py --workload kdyn256 --steps 1 --warmup 2 --no-cpu > $O/r2b_bench_kdyn256.json 2> $O/r2b_bench_kdyn256.err; echo "bench256 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:XFusedH -c 2 -o $O/prof_r2b_x256 -f python tools/ncu_kdyn.py 256 1 > $O/r2b_ncu256.log 2>&1; echo "ncu256 rc=$?"
tail -3 $O/r2b_tests.log; cat $O/r2b_sh23ens_ab.log; cat $O/r2b_time256.log
