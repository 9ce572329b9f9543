#!/bin/bash
# round-2 GPU session T (4 GPUs): last build - configs 3 and 4 at 4 GPUs (completes the 1/2/4/8 table)
cd "$(dirname "$0")/.."
O=gpurun_out
P=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $P --steps 5 --warmup 3 > $O/r2t_bench_kdyn128_p$P.json 2> $O/r2t_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $P --workload kdyn256 --steps 1 --warmup 2 > $O/r2t_bench_kdyn256_p$P.json 2> $O/r2t_bench_kdyn256_p$P.err; echo "bench256 rc=$?"
for f in kdyn128 kdyn256; do grep "^{" $O/r2t_bench_${f}_p$P.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'J', b.get('J'), 'mp', b.get('mp_parity_relerr'), b['config'].get('store'), b['config'].get('rho'), 'e2e', b['e2e']['value'], 'pair', (b.get('roofline_pair') or {}).get('frac'))
"; tail -2 $O/r2t_bench_${f}_p$P.err; done
