#!/bin/bash
# multi-GPU bench lines: bash profiles/scripts/multi.sh <n_gpus> [extra bench.py args...]   (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
TAG=${TAG:-cpm}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
cat gpurun_out/bench_${TAG}_${N}gpu.json
