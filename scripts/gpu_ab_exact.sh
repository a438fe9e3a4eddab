#!/bin/bash
# A/B of two library builds on the per-step trajectory path: ab/libahmc_base.so (previous build) vs the in-tree library.
mkdir -p gpurun_out
for rep in 1 2; do
  AHMC_B200_LIB=$PWD/ab/libahmc_base.so python scripts/ab_exact.py 2>&1 | tail -1 | tee -a gpurun_out/ab_exact.log
  python scripts/ab_exact.py 2>&1 | tail -1 | tee -a gpurun_out/ab_exact.log
done
