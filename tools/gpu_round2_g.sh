#!/bin/bash
# round-2 GPU session G (1 GPU): new SH23 kernels (grid-form snapshots), ensemble of optimisations
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs -k "sh23 or ensemble or sideout or trajectory or reference or abi or smoke or errors or stale or reproducible" > $O/r2g_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2g_tests.log
python tools/time_sh23ens.py 4096 > $O/r2g_sh23ens.log 2>&1
python tools/time_sh23ens.py 4736 >> $O/r2g_sh23ens.log 2>&1
python tools/time_sh23ens.py 2368 >> $O/r2g_sh23ens.log 2>&1
python bench.py --workload sh23ens --steps 5 --warmup 3 > $O/r2g_bench_sh23ens.json 2> $O/r2g_bench_sh23ens.err; echo "sh23ens rc=$?"
python bench.py --workload sh23 --steps 5 --warmup 3 > $O/r2g_bench_sh23.json 2> $O/r2g_bench_sh23.err; echo "sh23 rc=$?"
SMO_ENS_TOTAL=512 SMO_ENS_ITERS=20 timeout 900 python bench.py --workload sh23opt --steps 1 --warmup 0 > $O/r2g_bench_sh23opt512.json 2> $O/r2g_bench_sh23opt512.err; echo "sh23opt rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 $O/r2g_tests.log; cat $O/r2g_sh23ens.log; tail -c 1800 $O/r2g_bench_sh23opt512.json; tail -3 $O/r2g_bench_sh23opt512.err; tail -3 $O/r2g_smoke.log
