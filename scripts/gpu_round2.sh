#!/bin/bash
# GPU pass: full parity suite, e2e probe, BASELINE configs with timing breakdown, bench.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
timeout 900 python scripts/run_configs.py c1 c2 c3 2>&1 | grep -v Warn > gpurun_out/configs_1gpu.log
timeout 900 python scripts/run_configs.py c4 --scale 0.125 2>&1 | grep -v Warn >> gpurun_out/configs_1gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/e2e_probe.log; cat gpurun_out/configs_1gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
