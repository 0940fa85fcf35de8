#!/bin/bash
# Diagnostic GPU run: every test file under its own timeout, verbose and unbuffered, so that a hang names its test.
#   FILES="test_a test_b" bash scripts/gpu_diag.sh        (default: every *_gpu test file)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/diag.log
FILES=${FILES:-"test_stage0_gpu test_mt19937_gpu test_fof_gpu test_clean_gpu test_gpu_golden test_edge_cases_gpu test_giant_groups_gpu test_matching_gpu test_parity_sizes_gpu test_slabs_gpu test_transitive_gpu"}
for f in $FILES; do
  echo "=== $f" >> gpurun_out/diag.log
  timeout ${PER_FILE_TIMEOUT:-240} python -u -m pytest tests/$f.py -x -v -m gpu --timeout=${PER_TEST_TIMEOUT:-120} --timeout-method=thread >> gpurun_out/diag.log 2>&1
  echo "rc=$? $f" | tee -a gpurun_out/diag.log
done
grep -E "PASSED|FAILED|ERROR|rc=|Timeout|^E " gpurun_out/diag.log | tail -200 > gpurun_out/diag_summary.log
grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/diag_summary.log | head -40
