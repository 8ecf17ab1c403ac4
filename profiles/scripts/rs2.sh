#!/bin/bash
mkdir -p gpurun_out
SGB_LIBRARY=$PWD/build/variants/rs_mb4.so timeout 600 python -m pytest tests -m gpu -q -x -k "reset or spawn or shard or respawn or golden" 2>&1 | tail -40 > gpurun_out/rs2.log
cat gpurun_out/rs2.log
