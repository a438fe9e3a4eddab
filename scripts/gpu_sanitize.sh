#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the GPU tests that cover this round's new code paths
set -u
mkdir -p gpurun_out
: > gpurun_out/sanitizer_summary.log
SEL="user_target or nuts_on_a_user or pooled or stepsize or without_cached or autotune or small_host or nuts_transition_vs_oracle or dense_tile or every_register_layout or headline"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit: $?" | tee -a gpurun_out/sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3 | tee -a gpurun_out/sanitizer_summary.log
done
