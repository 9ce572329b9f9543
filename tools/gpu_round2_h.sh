#!/bin/bash
# round-2 GPU session H (2 GPUs): staggered chunk pipeline, config 4 (256^3, checkpointed) on 2 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out
P=${P:-2}
VARIANTS="1,1,0;1,2,1;1,2,2,1;1,3,2;1,4,2;2,4,2" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29534 tools/time_kdyn_mp.py 128 200 > $O/r2h_time_mp_p$P.log 2>&1; echo "time_mp rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $P --workload kdyn256 --steps 1 --warmup 1 > $O/r2h_bench_kdyn256_p$P.json 2> $O/r2h_bench_kdyn256_p$P.err; echo "bench256 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $P --steps 3 --warmup 3 > $O/r2h_bench_kdyn128_p$P.json 2> $O/r2h_bench_kdyn128_p$P.err; echo "bench128 rc=$?"
grep -v "Warning\|^\*\|OMP_NUM" $O/r2h_time_mp_p$P.log | tail -60; tail -c 2500 $O/r2h_bench_kdyn256_p$P.json; tail -5 $O/r2h_bench_kdyn256_p$P.err; tail -c 600 $O/r2h_bench_kdyn128_p$P.json
