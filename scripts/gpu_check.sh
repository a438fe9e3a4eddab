#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (+ reference arm), ncu launch list, BASELINE configs, K4 tile A/B.
# Outputs -> gpurun_out/.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh [prof]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python scripts/run_configs.py c1 c2 c3 c4 c4v c5 2>&1 | grep -v Warn > gpurun_out/configs_1gpu.log
timeout 300 python scripts/k4_ab.py > gpurun_out/k4_ab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
if [ "${1:-}" = "prof" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 2 -c 2 -f -o /tmp/k1 python scripts/profile_k1.py none > gpurun_out/prof.log 2>&1
  ncu -i /tmp/k1.ncu-rep --page raw --csv > gpurun_out/k1_headline_raw.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/k1.ncu-rep --page source --csv > gpurun_out/k1_headline_source.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/k1.ncu-rep --page details > gpurun_out/k1_headline_details.txt 2>> gpurun_out/prof.log
fi
if [ "${1:-}" = "hostlane" ]; then  # the zero-copy host-buffer launch (reads / writes pinned host memory directly)
  timeout 600 ncu --set full --clock-control none -k regex:leapfrog_kernel -s 6 -c 1 -f -o /tmp/k1h python scripts/profile_k1.py hostlane > gpurun_out/prof_hostlane.log 2>&1
  ncu -i /tmp/k1h.ncu-rep --page details > gpurun_out/k1_hostlane_details.txt 2>> gpurun_out/prof_hostlane.log
  ncu -i /tmp/k1h.ncu-rep --page raw --csv > gpurun_out/k1_hostlane_raw.csv 2>> gpurun_out/prof_hostlane.log
fi
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
cat gpurun_out/bench_ref.json; cat gpurun_out/configs_1gpu.log; cat gpurun_out/k4_ab.log
