#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/g1.log; : > $L
for so in build/variants/g1_mb1.so build/variants/g1_mb2.so; do
  for cfg in "8192 30 roundabout_2 20" "8192 30 interchange_2 17" "16384 30 cpm_mixed 20" "16384 20 cpm_entire 18"; do
    SGB_LIBRARY=$PWD/$so KB_WRITE_OBS=1 timeout 200 python profiles/kbench.py $cfg >> $L 2>&1
  done
done
cut -c1-215 $L
