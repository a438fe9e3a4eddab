#!/bin/bash
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "nutpie or welford_cov or adapt_cov or pipelined" 2>&1 | tail -8 > gpurun_out/pytest_sub.log
timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
tail -4 gpurun_out/pytest_sub.log; grep -v "^copies\|^N \|^H2D\|^D2H" gpurun_out/e2e_probe.log
