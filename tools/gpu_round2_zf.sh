#!/bin/bash
# session r2zf (1 GPU): GPU test suite and the default bench line of the final build
O=gpurun_out; mkdir -p $O
( time timeout 185 python -m pytest tests -m gpu -x -q -rs ) > $O/r2zf_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2zf_gpu_tests.log
timeout 50 python bench.py --steps 3 --warmup 3 > $O/r2zf_bench_kdyn128_n1.json 2> $O/r2zf_bench.err; echo "bench rc=$?"
tail -4 $O/r2zf_gpu_tests.log; cut -c1-600 $O/r2zf_bench_kdyn128_n1.json
