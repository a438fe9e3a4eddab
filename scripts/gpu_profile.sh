#!/bin/bash
# ncu captures (full set) of K1 (headline shape, cold L2; HBM-honest shape), K2, K3, K4; the .ncu-rep files stay on the box
# (too large for gpurun_out), their raw / source / details pages and a JSON summary come back.
set -u
mkdir -p gpurun_out
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
prof k4 dense_traj_kernel 1 1 k4
python scripts/ncu_summary.py gpurun_out/k1_headline_raw.csv gpurun_out/k1_honest_raw.csv gpurun_out/k2_raw.csv gpurun_out/k3_raw.csv gpurun_out/k4_raw.csv | tee gpurun_out/ncu_summary.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/prof.log
