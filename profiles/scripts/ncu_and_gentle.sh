#!/bin/bash
bash profiles/scripts/evidence.sh ncu | tail -3
L=gpurun_out/sweep_gentle_r2.log
: > $L
KB_WRITE_OBS=1 KB_GENTLE=1 timeout 300 python profiles/kbench.py 65536 40 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_GENTLE=1 KB_REW=ttc_sparse timeout 300 python profiles/kbench.py 65536 40 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_GENTLE=1 timeout 300 python profiles/kbench.py 65536 40 cpm_mixed 8 >> $L 2>&1
cat $L
