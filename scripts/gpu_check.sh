#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (+ reference arm), ncu launch list and K1 capture.
# Outputs -> gpurun_out/.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh [prof]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ "${1:-}" = "prof" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 2 -c 2 -f -o /tmp/k1 python scripts/profile_k1.py none > gpurun_out/prof.log 2>&1
  ncu -i /tmp/k1.ncu-rep --page raw --csv > gpurun_out/k1_headline_raw.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/k1.ncu-rep --page source --csv > gpurun_out/k1_headline_source.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/k1.ncu-rep --page details > gpurun_out/k1_headline_details.txt 2>> gpurun_out/prof.log
fi
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
cat gpurun_out/bench_ref.json
