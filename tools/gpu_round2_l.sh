#!/bin/bash
# round-2 GPU session L (1 GPU): final build - full suite, bench lines of every workload, ncu launch list + full capture
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/r2l_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2l_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > $O/r2l_bench_kdyn128.json 2> $O/r2l_bench_kdyn128.err; echo "bench128 rc=$?"
timeout 600 python bench.py --workload kdyn256 --steps 1 --warmup 2 --no-cpu > $O/r2l_bench_kdyn256.json 2> $O/r2l_bench_kdyn256.err; echo "bench256 rc=$?"
python bench.py --workload kdyn24 --steps 5 --warmup 3 > $O/r2l_bench_kdyn24.json 2> $O/r2l_bench_kdyn24.err; echo "kdyn24 rc=$?"
python bench.py --workload sh23ens --steps 5 --warmup 3 > $O/r2l_bench_sh23ens.json 2> $O/r2l_bench_sh23ens.err; echo "sh23ens rc=$?"
python bench.py --workload sh23 --steps 5 --warmup 3 > $O/r2l_bench_sh23.json 2> $O/r2l_bench_sh23.err; echo "sh23 rc=$?"
python bench.py --workload vec --steps 3 --warmup 3 > $O/r2l_bench_vec.json 2> $O/r2l_bench_vec.err; echo "vec rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2l_bench_reference_arm.json 2> $O/r2l_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3900 -c 400 --csv --log-file $O/r2l_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-graph > $O/r2l_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:6XFusedINS_3FacILi16ELi12EEE -c 4 -o $O/prof_r2l_x128 -f python tools/ncu_kdyn.py 128 2 > $O/r2l_ncu128.log 2>&1; echo "ncu128 rc=$?"
GRAPH=1 python tools/time_kdyn.py 128 50 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" > $O/r2l_time128.log
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" > $O/r2l_time256.log
tail -3 $O/r2l_tests.log; tail -2 $O/r2l_smoke.log; cat $O/r2l_time128.log $O/r2l_time256.log
