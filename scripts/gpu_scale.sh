#!/bin/bash
# 8-GPU evidence pass (one box): bench at N=8, the 8-rank NCCL exchange check, BASELINE configs C4 (three adaptation forms) and C5.
set -u
mkdir -p gpurun_out
N=${1:-8}
QUICK=${2:-}   # "quick": the bench line, C5 / C3 and the device-pooled C4 only
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
if [ "$QUICK" = "quick" ]; then
  : > gpurun_out/nccl_check_${N}gpu.log
  timeout 600 $TR --master-port 29543 scripts/run_configs.py c4d --iters 1250 2>&1 | grep -v Warn > gpurun_out/configs_${N}gpu_1250.log
else
  timeout 300 $TR --master-port 29542 scripts/nccl_exchange_check.py > gpurun_out/nccl_check_${N}gpu.log 2>&1
  timeout 900 $TR --master-port 29543 scripts/run_configs.py c4d c4 c4v --iters 1250 2>&1 | grep -v Warn > gpurun_out/configs_${N}gpu_1250.log
fi
timeout 600 $TR --master-port 29544 scripts/run_configs.py c5 c3 2>&1 | grep -v Warn > gpurun_out/configs_${N}gpu.log
cut -c1-1200 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/nccl_check_${N}gpu.log; cat gpurun_out/configs_${N}gpu_1250.log gpurun_out/configs_${N}gpu.log
