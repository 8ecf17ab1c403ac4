#!/bin/bash
# ncu --set full (with source) of the step / refresh kernels of the in-tree library, then the variant timings
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:env_step_kernel -s 8 -c 2 -f -o gpurun_out/cur_r2 python profiles/prof_step.py > gpurun_out/ncu_cur.log 2>&1
tail -3 gpurun_out/ncu_cur.log
bash profiles/scripts/variants.sh check
