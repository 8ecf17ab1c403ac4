#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/g2b.log; : > $L
for cfg in "8192 40 roundabout_2 12" "8192 40 on_ramp_2_multilane 12" "32768 30 roundabout_2 12" "32768 30 cpm_entire 15" "65536 30 cpm_entire 8"; do
  KB_WRITE_OBS=1 timeout 200 python profiles/kbench.py $cfg >> $L 2>&1
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $L
cut -c1-215 $L
