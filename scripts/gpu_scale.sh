#!/bin/bash
# multi-GPU pass (run under `gpurun --gpus 8`): weak-scaling bench at N=8 (+ optionally 2,4) and the sharded configs
set -u
mkdir -p gpurun_out
python advancedhmc.jl_b200/build.py > gpurun_out/build.log 2>&1
for n in ${SCALE_NS:-8}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-extras > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${n}gpu.json'));print($n, d['value'], d['ms_per_step'], d['e2e'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
    scripts/run_configs.py c4 c4v c5 2> gpurun_out/configs_8gpu.err | grep '^{' > gpurun_out/configs_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 \
    scripts/run_configs.py c4 c4v --iters 1250 2>> gpurun_out/configs_8gpu.err | grep '^{' >> gpurun_out/configs_8gpu.log
cat gpurun_out/configs_8gpu.log; tail -3 gpurun_out/configs_8gpu.err
