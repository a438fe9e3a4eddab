#!/bin/bash
# Round-2 starter: validate and time the staged NUTS variants against the default build.
#   1. HERE, before gpurun (the variant .so files travel with the snapshot, ~50 MB of push each):
#        scripts/build_variants.sh fastdraw fastdraw+fulltile fastdraw+fulltile+altlayout1 fastdraw+altlayout2 \
#            fastdraw+fulltile+reloadcoef+minb4
#   2. gpurun --timeout 2400 -- 'bash scripts/gpu_fastdraw_ab.sh'
# For every variant found: the GPU parity tests that touch NUTS run with AHMC_B200_LIB pointing at it, then
# scripts/nuts_ab.py times it.  Outputs -> gpurun_out/ab_<tag>_*.  Make a knob the default (flip the macro in
# ahmc_nuts_kernel.cuh) only if its parity run is green and nuts_ab shows it faster.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/nuts_ab.py > gpurun_out/ab_default.jsonl 2>&1
echo "--- default"; grep '^{' gpurun_out/ab_default.jsonl
for W in advancedhmc.jl_b200/_variants/libahmc_b200_*.so; do
  [ -f "$W" ] || { echo "no variants built: scripts/build_variants.sh fastdraw ..."; exit 1; }
  tag=$(basename "$W" .so); tag=${tag#libahmc_b200_}
  AHMC_B200_LIB=$PWD/$W timeout 900 python -m pytest tests -m gpu -q -k "nuts or in_launch or vectorised or mp50 or c3 or c4 or c5" 2>&1 | tail -3 > gpurun_out/ab_${tag}_pytest.log
  AHMC_B200_LIB=$PWD/$W timeout 300 python scripts/nuts_ab.py > gpurun_out/ab_${tag}.jsonl 2>&1
  echo "--- $tag"; tail -1 gpurun_out/ab_${tag}_pytest.log; grep '^{' gpurun_out/ab_${tag}.jsonl
done
