#!/bin/bash
# round-2 GPU session I (8 GPUs): parity at P = 8, scaling of configs 3 / 4 / 5
cd "$(dirname "$0")/.."
O=gpurun_out
P=${P:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29541 tests/mp_parity.py > $O/r2i_mp_parity_p$P.log 2>&1; echo "mp_parity rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $P --steps 5 --warmup 3 > $O/r2i_bench_kdyn128_p$P.json 2> $O/r2i_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
VARIANTS="1,1,0,1;1,3,1;1,3,2" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29543 tools/time_kdyn_mp.py 128 200 > $O/r2i_time_mp_p$P.log 2>&1; echo "time_mp rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $P --workload kdyn256 --steps 2 --warmup 2 > $O/r2i_bench_kdyn256_p$P.json 2> $O/r2i_bench_kdyn256_p$P.err; echo "bench256 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $P --workload sh23ens --steps 5 --warmup 3 > $O/r2i_bench_sh23ens_p$P.json 2> $O/r2i_bench_sh23ens_p$P.err; echo "sh23ens rc=$?"
grep "MP_PARITY\|max rel" $O/r2i_mp_parity_p$P.log | tail -16; grep -v "Warning\|^\*\|OMP_NUM" $O/r2i_time_mp_p$P.log | tail -40
for f in kdyn128 kdyn256 sh23ens; do grep "^{" $O/r2i_bench_${f}_p$P.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'J', b.get('J'), 'mp', b.get('mp_parity_relerr'), b['config'].get('store'), 'e2e', b['e2e']['value'], 'roofline', b['roofline']['launch_ms'], b['roofline']['frac'], 'pair', (b.get('roofline_pair') or {}).get('frac'))
"; tail -2 $O/r2i_bench_${f}_p$P.err; done
