#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/rs4.log; : > $L
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 32768 30 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 32768 30 cpm_entire 15 >> $L 2>&1
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/prof_variants.py cpm_entire 8 20000 > gpurun_out/sanitizer_${tool}_subwarp_n8.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/prof_variants.py roundabout_2 12 20000 > gpurun_out/sanitizer_${tool}_subwarp_n12.log 2>&1
  grep -E "SUMMARY" gpurun_out/sanitizer_${tool}_subwarp_n8.log gpurun_out/sanitizer_${tool}_subwarp_n12.log >> $L
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $L
cut -c1-215 $L
