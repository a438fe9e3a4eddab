#!/bin/bash
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/run_configs.py c4 c4v --scale 0.125 2>&1 | grep -v Warn > gpurun_out/configs_c4.log
timeout 600 python scripts/run_configs.py c4 c4v --scale 0.125 --iters 1000 2>&1 | grep -v Warn >> gpurun_out/configs_c4.log
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/configs_c4.log
