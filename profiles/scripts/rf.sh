#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/rf.log; : > $L
for so in build/variants/rf_sw*.so; do
  SGB_LIBRARY=$PWD/$so KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 >> $L 2>&1
  SGB_LIBRARY=$PWD/$so KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 >> $L 2>&1
done
cut -c1-215 $L
