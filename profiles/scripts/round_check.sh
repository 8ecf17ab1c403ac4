#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "TIES|passed|failed|Error|error|assert" | tail -80 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -60
bash profiles/scripts/variants.sh check | tail -6
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_smoke.log 2>&1
tail -3 gpurun_out/sanitizer_racecheck_smoke.log
