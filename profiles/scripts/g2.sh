#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/g2.log; : > $L
for so in build/variants/g2_mb1.so build/variants/g2_mb2.so; do
  for cfg in "8192 40 roundabout_2 12" "8192 40 on_ramp_2_multilane 12" "32768 30 roundabout_2 12" "32768 30 cpm_entire 15" "16384 30 cpm_mixed 12"; do
    SGB_LIBRARY=$PWD/$so KB_WRITE_OBS=1 timeout 200 python profiles/kbench.py $cfg >> $L 2>&1
  done
done
cut -c1-215 $L
