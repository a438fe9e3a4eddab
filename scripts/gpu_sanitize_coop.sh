#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the GPU tests that run the cooperative dense NUTS products
# (bulk copies + mbarriers + DMMA, blocked triangular solve) and the tempered transitions
set -u
mkdir -p gpurun_out
: > gpurun_out/sanitizer_coop_summary.log
SEL="tempered_leapfrog_inside or nuts_transition_vs_oracle"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_coop_$tool.log 2>&1
  echo "$tool exit: $?" | tee -a gpurun_out/sanitizer_coop_summary.log
  grep -E "ERROR SUMMARY|passed|failed|Race reported|hazard" gpurun_out/sanitizer_coop_$tool.log | sort | uniq -c | sort -rn | head -8 | tee -a gpurun_out/sanitizer_coop_summary.log
done
