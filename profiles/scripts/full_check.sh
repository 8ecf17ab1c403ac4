#!/bin/bash
# GPU tests + default bench line (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -3 gpurun_out/bench_1gpu.err
cat gpurun_out/bench_1gpu.json
