#!/bin/bash
# round-2 GPU session W (1 GPU): the A/B switches of the last build still give oracle parity (grid accumulation off; TMA copies off)
cd "$(dirname "$0")/.."
O=gpurun_out
for opts in "7=0" "8=0,9=0" "7=0,8=0,9=0"; do
  SMO_KDYN_OPTS="$opts" python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kdyn_f_gradf or checkpointed or integrated or graph_replay or large_grids" > $O/r2w_tests_$(echo $opts | tr ',=' '__').log 2>&1; echo "tests($opts) rc=$?"
  tail -1 $O/r2w_tests_$(echo $opts | tr ',=' '__').log
done
