#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/quick2.log; : > $L
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 40 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 40 >> $L 2>&1
KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 40 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 KB_LAYOUT=is_obs_steering=1,is_observe_ref_path_other_agents=1 timeout 120 python profiles/kbench.py 65536 40 >> $L 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L | cut -c1-230
