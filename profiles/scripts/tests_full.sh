#!/bin/bash
# Full GPU suite (no -x so every test reports) + default 1-GPU bench line + reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_full.log
tail -15 gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -3 gpurun_out/bench_1gpu.err
cat gpurun_out/bench_1gpu.json
