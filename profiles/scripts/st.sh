#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/st.log; : > $L
for so in build/variants/st_full.so build/variants/st_twice.so; do
  for B in 4736 9472 65536; do
    SGB_LIBRARY=$PWD/$so timeout 120 python profiles/kbench.py $B 40 >> $L 2>&1
  done
done
cut -c1-215 $L
