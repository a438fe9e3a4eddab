#!/bin/bash
# ncu captures (full set) for K1 (headline + honest), K2, K3; reports are exported to CSV on the box
# (raw + source pages) because the .ncu-rep files exceed the 64 MiB gpurun_out budget.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
prof() {  # name kernel-regex skip count args...
  local name=$1 kre=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c $cnt -f -o /tmp/$name python scripts/profile_k1.py "$@" >> gpurun_out/prof.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>> gpurun_out/prof.log
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>> gpurun_out/prof.log
}
: > gpurun_out/prof.log
prof k1_headline leapfrog_kernel 2 2 none
prof k1_honest leapfrog_kernel 5 1 honest
prof k2 hmc_kernel 1 1 k2
prof k3 nuts_kernel 1 1 k3
ls -la gpurun_out; tail -3 gpurun_out/prof.log
