#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over a representative subset of the GPU tests
set -u
mkdir -p gpurun_out
SEL="golden or every_register_layout or nuts_transition_vs_oracle or hmc_transition_vs_oracle or callback_target or multi_transition or adapt_summary or nonfinite"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit: $?" | tee -a gpurun_out/sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3 | tee -a gpurun_out/sanitizer_summary.log
done
