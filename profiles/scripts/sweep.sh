#!/bin/bash
# SURVEY.md §8(d) input list: kernel ms per step over reward methods / action distributions / batch sizes / maps / layouts
mkdir -p gpurun_out
L=gpurun_out/sweep_r2.log
: > $L
K="timeout 200 python profiles/kbench.py"
KB_WRITE_OBS=1 $K 65536 30 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_REW=ttc_sparse $K 65536 30 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_GENTLE=1 $K 65536 40 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_GENTLE=1 KB_REW=ttc_sparse $K 65536 40 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 $K 8192 30 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 $K 4736 30 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 $K 9472 30 cpm_entire 8 >> $L 2>&1
KB_WRITE_OBS=1 $K 65536 30 cpm_mixed 8 >> $L 2>&1
KB_WRITE_OBS=1 KB_REW=ttc_sparse $K 32768 30 cpm_entire 15 >> $L 2>&1
KB_WRITE_OBS=1 $K 8192 30 on_ramp_2_multilane 12 >> $L 2>&1
KB_WRITE_OBS=1 $K 8192 30 roundabout_2 12 >> $L 2>&1
KB_WRITE_OBS=1 $K 32768 30 roundabout_2 12 >> $L 2>&1
for lay in is_obs_steering=1,is_observe_ref_path_other_agents=1 is_ego_view=0 is_observe_distance_to_boundaries=0 is_use_mtv_distance=1 is_apply_mask=1 is_obs_noise=1; do
  KB_WRITE_OBS=1 KB_LAYOUT=$lay $K 65536 30 cpm_entire 8 >> $L 2>&1
done
cat $L
