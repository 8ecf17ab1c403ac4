#!/bin/bash
# Development helper: time (and bit-check) every library variant under build/variants/ at the headline shape.
#   gpurun -- 'bash profiles/scripts/variants.sh [check]'
mkdir -p gpurun_out
: > gpurun_out/variants.log
for so in build/variants/*.so; do
  SGB_LIBRARY=$PWD/$so timeout 120 python profiles/kbench.py 65536 30 >> gpurun_out/variants.log 2>&1
  if [ "$1" = "check" ] && [ "$(basename $so)" != "base.so" ]; then
    SGB_LIBRARY=$PWD/$so timeout 300 python profiles/exact_check.py 32768 12 >> gpurun_out/variants.log 2>&1
  fi
done
cat gpurun_out/variants.log
