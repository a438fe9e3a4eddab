#!/bin/bash
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/k4_ab.py > gpurun_out/k4_ab.log 2>&1
timeout 600 python scripts/run_configs.py c5 c2 2>&1 | grep -v Warn > gpurun_out/configs_c5.log
AHMC_DENSE_TILE=16x2 timeout 600 python scripts/run_configs.py c2 2>&1 | grep -v Warn >> gpurun_out/configs_c5.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/k4_ab.log; cat gpurun_out/configs_c5.log
