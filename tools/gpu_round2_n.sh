#!/bin/bash
# round-2 GPU session N (2 GPUs): final build on 2 GPUs - parity of every transport, configs 3 and 4
cd "$(dirname "$0")/.."
O=gpurun_out
P=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29561 tests/mp_parity.py > $O/r2n_mp_parity_p$P.log 2>&1; echo "mp_parity rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $P --steps 3 --warmup 3 > $O/r2n_bench_kdyn128_p$P.json 2> $O/r2n_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $P --workload kdyn256 --steps 1 --warmup 1 > $O/r2n_bench_kdyn256_p$P.json 2> $O/r2n_bench_kdyn256_p$P.err; echo "bench256 rc=$?"
for i in 1 2 3; do python tools/time_sh23ens.py 4096 >> $O/r2n_sh23_spread.log 2>&1; done
grep "MP_PARITY\|max rel" $O/r2n_mp_parity_p$P.log | tail -5; cat $O/r2n_sh23_spread.log
for f in kdyn128 kdyn256; do grep "^{" $O/r2n_bench_${f}_p$P.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'J', b.get('J'), 'mp', b.get('mp_parity_relerr'), b['config'].get('store'), 'e2e', b['e2e']['value'], 'pair', (b.get('roofline_pair') or {}).get('frac'))
"; tail -2 $O/r2n_bench_${f}_p$P.err; done
