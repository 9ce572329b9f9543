#!/bin/bash
# round-2 GPU session X (1 GPU): verbose log of the history-parity tests (printed worst relative errors)
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_history.py -m gpu -v -s -W ignore > $O/r2x_history_tests_verbose.log 2>&1; echo "rc=$?" | tee -a $O/r2x_history_tests_verbose.log
grep -v "^$" $O/r2x_history_tests_verbose.log | tail -30
