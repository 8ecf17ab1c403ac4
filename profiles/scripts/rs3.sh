#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/rs3.log; : > $L
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 on_ramp_2_multilane 12 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 30 cpm_mixed 8 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 30 roundabout_2 20 >> $L 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $L
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/prof_variants.py roundabout_2 12 64 > gpurun_out/sanitizer_${tool}_n12.log 2>&1
  tail -1 gpurun_out/sanitizer_${tool}_smoke.log gpurun_out/sanitizer_${tool}_n12.log >> $L
done
cut -c1-215 $L
