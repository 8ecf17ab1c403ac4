#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1
cat gpurun_out/e2e_sweep.log
