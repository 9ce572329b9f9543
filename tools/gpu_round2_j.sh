#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
P=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $P --steps 3 --warmup 3 > $O/r2j_bench_kdyn128_p$P.json 2> $O/r2j_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $P --workload sh23ens --steps 5 --warmup 3 > $O/r2j_bench_sh23ens_p$P.json 2> $O/r2j_bench_sh23ens_p$P.err; echo "sh23ens rc=$?"
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "two" > $O/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2j_tests.log
for f in kdyn128 sh23ens; do grep "^{" $O/r2j_bench_${f}_p$P.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'J', b.get('J'), 'mp', b.get('mp_parity_relerr'), 'e2e', b['e2e']['value'])
"; tail -2 $O/r2j_bench_${f}_p$P.err; done
