#!/bin/bash
# sanitizer over the code added this session + ncu capture of K4
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
SEL="variants or in_launch or vectorised or pipelined or adapt_cov or nutpie or welford_cov or multinomial"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck_new.log 2>&1
echo "memcheck exit: $?" | tee gpurun_out/sanitizer_new_summary.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_new.log | tail -3 | tee -a gpurun_out/sanitizer_new_summary.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "adapt_cov or in_launch or variants_vs_oracle" > gpurun_out/sanitizer_racecheck_new.log 2>&1
echo "racecheck exit: $?" | tee -a gpurun_out/sanitizer_new_summary.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_new.log | tail -3 | tee -a gpurun_out/sanitizer_new_summary.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_traj_kernel -s 1 -c 1 -f -o /tmp/k4 python scripts/profile_k1.py k4 > gpurun_out/prof_k4.log 2>&1
ncu -i /tmp/k4.ncu-rep --page details > gpurun_out/k4_details.txt 2>> gpurun_out/prof_k4.log
ncu -i /tmp/k4.ncu-rep --page source --csv > gpurun_out/k4_source.csv 2>> gpurun_out/prof_k4.log
tail -3 gpurun_out/prof_k4.log
