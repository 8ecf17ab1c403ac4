#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./profiles/micro/f32x2_bench > gpurun_out/f32x2_bench.log 2>&1
cat gpurun_out/f32x2_bench.log
bash profiles/scripts/variants.sh check
