#!/bin/bash
# round-2 GPU session P (8 GPUs): final build - configs 3 / 4 / 5 (as optimisations), A/B of the relaxed hand-shake build
cd "$(dirname "$0")/.."
O=gpurun_out
P=${P:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $P --steps 5 --warmup 3 > $O/r2p_bench_kdyn128_p$P.json 2> $O/r2p_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
SMO_B200_LIB=$PWD/build/libsmo_relaxed.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $P --steps 5 --warmup 3 > $O/r2p_bench_kdyn128_relaxed_p$P.json 2> $O/r2p_bench_kdyn128_relaxed_p$P.err; echo "bench128 relaxed rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $P --workload kdyn256 --steps 2 --warmup 2 > $O/r2p_bench_kdyn256_p$P.json 2> $O/r2p_bench_kdyn256_p$P.err; echo "bench256 rc=$?"
SMO_ENS_TOTAL=4096 SMO_ENS_ITERS=20 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus $P --workload sh23opt --steps 1 --warmup 0 > $O/r2p_bench_sh23opt_p$P.json 2> $O/r2p_bench_sh23opt_p$P.err; echo "sh23opt rc=$?"
for f in kdyn128 kdyn128_relaxed kdyn256 sh23opt; do grep "^{" $O/r2p_bench_${f}_p$P.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$f', 'ms/step', b['ms_per_step'], 'value', b['value'], 'J', b.get('J'), 'mp', b.get('mp_parity_relerr'), b['config'].get('store'), 'e2e', b['e2e']['value'], 'pair', (b.get('roofline_pair') or {}).get('frac'), b.get('ensemble'))
"; tail -2 $O/r2p_bench_${f}_p$P.err; done
