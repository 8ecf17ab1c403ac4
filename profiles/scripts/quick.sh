#!/bin/bash
# Development loop: headline-shape kernel timing + bitwise pruned == exhaustive check (+ optional pytest filter)
#   gpurun -- 'bash profiles/scripts/quick.sh ["pytest -k expression"]'
mkdir -p gpurun_out
L=gpurun_out/quick.log
: > $L
timeout 120 python profiles/kbench.py 65536 40 >> $L 2>&1
timeout 120 python profiles/kbench.py 8192 40 roundabout_2 12 >> $L 2>&1
timeout 300 python profiles/exact_check.py 32768 12 >> $L 2>&1
if [ -n "$1" ]; then timeout 1200 python -m pytest tests -m gpu -q -x -k "$1" 2>&1 | tail -4 >> $L; fi
cat $L
