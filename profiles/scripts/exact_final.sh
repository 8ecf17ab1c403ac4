#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exact_check_r2.log
: > $L
timeout 600 python profiles/exact_check.py 65536 40 cpm_entire 8 >> $L 2>&1
timeout 600 python profiles/exact_check.py 16384 60 cpm_mixed 6 >> $L 2>&1
timeout 600 python profiles/exact_check.py 16384 40 roundabout_2 12 >> $L 2>&1
timeout 600 python profiles/exact_check.py 8192 40 on_ramp_2_multilane 12 >> $L 2>&1
timeout 600 python profiles/exact_check.py 8192 30 cpm_entire 15 >> $L 2>&1
timeout 600 python profiles/exact_check.py 8192 30 roundabout_2 20 >> $L 2>&1
timeout 600 python profiles/exact_check.py 4096 30 cpm_entire 18 >> $L 2>&1
cat $L
