#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/epw.log; : > $L
for e in 1 2 3 4 6; do echo "epw $e" >> $L; SGB_RESET_EPW=$e KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 40 >> $L 2>&1; done
for e in 1 2 3; do echo "epw $e roundabout" >> $L; SGB_RESET_EPW=$e KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 40 roundabout_2 12 >> $L 2>&1; done
cat $L | cut -c1-220
