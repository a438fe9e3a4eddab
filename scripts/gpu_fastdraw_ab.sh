#!/bin/bash
# Round-2 starter: validate and time the staged NUTS variant (-DAHMC_NUTS_FASTDRAW=1) against the default build.
#   1. scripts/build_variants.sh fastdraw [altlayout fastdraw_altlayout]   (run HERE before gpurun: the variant .so files
#      travel with the snapshot; each is ~50 MB of push)
#   2. gpurun --timeout 1800 -- 'bash scripts/gpu_fastdraw_ab.sh'
# Outputs -> gpurun_out/fastdraw_*.  Make it the default (flip the macro in ahmc_nuts_kernel.cuh) only if the
# parity tests pass with the variant and nuts_ab shows it faster.
set -u
mkdir -p gpurun_out
V=advancedhmc.jl_b200/_variants/libahmc_b200_fastdraw.so
[ -f "$V" ] || { echo "build the variant first: scripts/build_variants.sh fastdraw"; exit 1; }
AHMC_B200_LIB=$PWD/$V timeout 900 python -m pytest tests -m gpu -q -k "nuts or in_launch or vectorised or mp50 or c3 or c4 or c5" 2>&1 | tail -15 > gpurun_out/fastdraw_pytest.log
echo "variant pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/fastdraw_pytest.log
timeout 300 python scripts/nuts_ab.py > gpurun_out/fastdraw_ab_default.jsonl 2>&1
AHMC_B200_LIB=$PWD/$V timeout 300 python scripts/nuts_ab.py > gpurun_out/fastdraw_ab_variant.jsonl 2>&1
tail -4 gpurun_out/fastdraw_pytest.log; echo "--- default"; grep '^{' gpurun_out/fastdraw_ab_default.jsonl; echo "--- fastdraw"; grep '^{' gpurun_out/fastdraw_ab_variant.jsonl
for tag in altlayout fastdraw_altlayout fastdraw_altlayout2; do   # optional: two chains per warp for 32 < D <= 128
  W=advancedhmc.jl_b200/_variants/libahmc_b200_$tag.so
  [ -f "$W" ] || continue
  AHMC_B200_LIB=$PWD/$W timeout 900 python -m pytest tests -m gpu -q -k "nuts or in_launch or mp50 or c3 or c4" 2>&1 | tail -3 > gpurun_out/${tag}_pytest.log
  AHMC_B200_LIB=$PWD/$W timeout 300 python scripts/nuts_ab.py > gpurun_out/${tag}_ab.jsonl 2>&1
  echo "--- $tag"; tail -1 gpurun_out/${tag}_pytest.log; grep '^{' gpurun_out/${tag}_ab.jsonl
done
