#!/bin/bash
# GPU pass: host-lane parity tests, e2e probe (transport sweep).  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "host or pipelined or Host or headline" 2>&1 | tail -8 > gpurun_out/pytest_host.log
timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
tail -4 gpurun_out/pytest_host.log; cat gpurun_out/e2e_probe.log
