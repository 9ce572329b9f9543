#!/bin/bash
# round-2 GPU session V (1 GPU): full GPU suite of the last commit
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -rs --durations=6 > $O/r2v_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2v_tests.log
tail -14 $O/r2v_tests.log
