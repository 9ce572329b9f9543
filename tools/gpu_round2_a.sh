#!/bin/bash
# round-2 GPU session A (1 GPU): tests, bench lines of every workload, per-kernel split
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs --durations=8 > $O/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2a_tests.log
python bench.py --steps 3 --warmup 3 > $O/r2a_bench_kdyn128.json 2> $O/r2a_bench_kdyn128.err; echo "bench128 rc=$?"
python bench.py --workload vec --steps 3 --warmup 3 > $O/r2a_bench_vec.json 2> $O/r2a_bench_vec.err; echo "vec rc=$?"
python bench.py --workload sh23 --steps 5 --warmup 3 > $O/r2a_bench_sh23.json 2> $O/r2a_bench_sh23.err; echo "sh23 rc=$?"
python bench.py --workload sh23ens --steps 5 --warmup 3 > $O/r2a_bench_sh23ens.json 2> $O/r2a_bench_sh23ens.err; echo "sh23ens rc=$?"
python bench.py --workload kdyn24 --steps 5 --warmup 3 > $O/r2a_bench_kdyn24.json 2> $O/r2a_bench_kdyn24.err; echo "kdyn24 rc=$?"
GRAPH=1 python tools/time_kdyn.py 128 50 > $O/r2a_time128.log 2>&1; echo "time128 rc=$?"
GRAPH=1 python tools/time_kdyn.py 256 10 > $O/r2a_time256.log 2>&1; echo "time256 rc=$?"
timeout 600 python bench.py --workload kdyn256 --steps 1 --warmup 2 --no-cpu > $O/r2a_bench_kdyn256.json 2> $O/r2a_bench_kdyn256.err; echo "bench256 rc=$?"
tail -3 $O/r2a_tests.log
