#!/bin/bash
# round-2 GPU session S (1 GPU): last build of the round - full suite, smoke, headline + config 4
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/r2s_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > $O/r2s_bench_kdyn128.json 2> $O/r2s_bench_kdyn128.err; echo "bench128 rc=$?"
timeout 600 python bench.py --workload kdyn256 --steps 1 --warmup 2 --no-cpu > $O/r2s_bench_kdyn256.json 2> $O/r2s_bench_kdyn256.err; echo "bench256 rc=$?"
GRAPH=1 python tools/time_kdyn.py 256 10 2>&1 | grep "N=\|x-adj\|x-fwd\|z-step\|y-pass" > $O/r2s_time256.log
tail -3 $O/r2s_tests.log; tail -2 $O/r2s_smoke.log; cat $O/r2s_time256.log
for f in kdyn128 kdyn256; do python -c "
import json
b = json.load(open('$O/r2s_bench_$f.json')); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'e2e', b['e2e']['value'], 'roofline', b['roofline']['launch_ms'], b['roofline']['frac'], 'pair', (b.get('roofline_pair') or {}).get('frac'), b['clocks'])
"; done
