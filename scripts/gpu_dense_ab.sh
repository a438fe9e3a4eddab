#!/bin/bash
# Round-2 starter for K4 (the dense tile kernel): validate and time the staged pipeline variants against the default build.
#   1. HERE, before gpurun:  scripts/build_variants.sh densepad denserel densepad+denserel densepad+denserel+dense3
#   2. gpurun --timeout 1500 -- 'bash scripts/gpu_dense_ab.sh'
# For every dense* variant found: the GPU parity tests that reach K4 run with AHMC_B200_LIB pointing at it, then
# scripts/k4_ab.py times it (4096 and 16384 chains x D = 128, L = 32; both tile shapes).  Outputs -> gpurun_out/k4ab_<tag>*.
# Make a knob the default (flip the macro in ahmc_dense.cu / ahmc_kernels.cuh) only if its parity run is green and
# k4_ab shows it faster.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/k4_ab.py > gpurun_out/k4ab_default.log 2>&1
echo "--- default"; grep 'ms/trajectory' gpurun_out/k4ab_default.log
for W in advancedhmc.jl_b200/_variants/libahmc_b200_dense*.so; do
  [ -f "$W" ] || { echo "no variants built: scripts/build_variants.sh densepad ..."; exit 1; }
  tag=$(basename "$W" .so); tag=${tag#libahmc_b200_}
  AHMC_B200_LIB=$PWD/$W timeout 600 python -m pytest tests -m gpu -q -k "dense or k4 or c2 or c5 or tile" 2>&1 | tail -3 > gpurun_out/k4ab_${tag}_pytest.log
  AHMC_B200_LIB=$PWD/$W timeout 300 python scripts/k4_ab.py > gpurun_out/k4ab_${tag}.log 2>&1
  echo "--- $tag"; tail -1 gpurun_out/k4ab_${tag}_pytest.log; grep 'ms/trajectory' gpurun_out/k4ab_${tag}.log
done
