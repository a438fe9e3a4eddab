#!/bin/bash
# ncu full capture of the COOP NUTS kernel on the C5 shape (one transition, N chains), pages + source-top back in gpurun_out/
set -u
mkdir -p gpurun_out
N=${1:-2048}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -s 1 -c 1 -f -o /tmp/c5 python scripts/profile_c5.py $N > gpurun_out/c5_prof.log 2>&1
ncu -i /tmp/c5.ncu-rep --page raw --csv > gpurun_out/c5_raw.csv 2>> gpurun_out/c5_prof.log
ncu -i /tmp/c5.ncu-rep --page source --csv > gpurun_out/c5_source.csv 2>> gpurun_out/c5_prof.log
ncu -i /tmp/c5.ncu-rep --page details > gpurun_out/c5_details.txt 2>> gpurun_out/c5_prof.log
python scripts/ncu_source_top.py gpurun_out/c5_source.csv 0 70 > gpurun_out/c5_source_top.txt 2>&1; python scripts/ncu_source_top.py gpurun_out/c5_source.csv 1 70 > gpurun_out/c5_srcline_top.txt 2>&1
tail -3 gpurun_out/c5_prof.log
python scripts/run_configs.py c5 --scale 8 2>&1 | tail -3
