#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/rs.log; : > $L
for so in build/variants/rs_mb6.so build/variants/rs_mb4.so; do
  for e in 3 8 16 32; do
    echo "== $so epw $e" >> $L
    SGB_LIBRARY=$PWD/$so SGB_RESET_EPW=$e timeout 120 python profiles/kbench.py 65536 30 >> $L 2>&1
  done
done
SGB_LIBRARY=$PWD/build/variants/rs_mb4.so timeout 600 python -m pytest tests -m gpu -q -x -k "reset or spawn or shard or respawn or golden" 2>&1 | tail -3 >> $L
cut -c1-215 $L
