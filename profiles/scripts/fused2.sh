#!/bin/bash
# fused GAE + all-gather on N GPUs: correctness (both store flavours), then bench lines with each flavour
#   gpurun --gpus N -- 'bash profiles/scripts/fused2.sh N'
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/fused_${N}gpu.log
: > $L
for mc in 0 1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/tools/fused_gather_check.py $mc 2>&1 | grep -E "FUSED|Error|error|assert" >> $L
done
for mc in ${MCS:-0 1}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --multicast $mc > gpurun_out/bench_fused_mc${mc}_${N}gpu.json 2> gpurun_out/bench_fused_mc${mc}_${N}gpu.err
  tail -2 gpurun_out/bench_fused_mc${mc}_${N}gpu.err >> $L
  python - >> $L <<PY
import json
for l in open("gpurun_out/bench_fused_mc${mc}_${N}gpu.json"):
    if l.startswith("{"):
        r = json.loads(l)["rollout"]; print("multicast=$mc", r["breakdown_ms"], r.get("nccl_two_step_ms"), r.get("fused_equals_gae_plus_nccl_all_gather"), "%.4e" % r["value"])
PY
done
cat $L
