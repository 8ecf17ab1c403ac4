#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "reset_draws or sets_cpm_mixed or mtv_distance_changes" 2>&1 | grep -v "^$" | head -400 > gpurun_out/pytest_failed.log
tail -5 gpurun_out/pytest_failed.log
