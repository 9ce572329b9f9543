#!/bin/bash
# round-2 GPU session O (1 GPU): final build (TMA tensor copies, 2-column half-length pass, carve-out hint, slab copies) - full suite + headline
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/r2o_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2o_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2o_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > $O/r2o_bench_kdyn128.json 2> $O/r2o_bench_kdyn128.err; echo "bench128 rc=$?"
timeout 600 python bench.py --workload kdyn256 --steps 1 --warmup 2 --no-cpu > $O/r2o_bench_kdyn256.json 2> $O/r2o_bench_kdyn256.err; echo "bench256 rc=$?"
python bench.py --workload sh23ens --steps 5 --warmup 3 > $O/r2o_bench_sh23ens.json 2> $O/r2o_bench_sh23ens.err; echo "sh23ens rc=$?"
tail -3 $O/r2o_tests.log; tail -2 $O/r2o_smoke.log
for f in kdyn128 kdyn256 sh23ens; do python -c "
import json
b = json.load(open('$O/r2o_bench_$f.json')); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'e2e', b['e2e']['value'], 'roofline', b['roofline']['launch_ms'], b['roofline']['frac'], 'pair', (b.get('roofline_pair') or {}).get('frac'), b.get('ms_forward'), b.get('ms_adjoint'))
"; done
