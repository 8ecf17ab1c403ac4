#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/rs5.log; : > $L
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 32768 30 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 32768 30 cpm_entire 15 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 cpm_mixed 8 >> $L 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $L
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
cut -c1-215 $L
