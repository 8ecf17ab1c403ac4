#!/bin/bash
# Round evidence: ncu --set full of every kernel family, launch list of the bench, sanitizer logs.
#   gpurun --timeout 2400 -- 'bash profiles/scripts/evidence.sh [ncu|san|all]'
W=${1:-all}
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
if [ "$W" = ncu ] || [ "$W" = all ]; then
  timeout 400 $NCU -k regex:env_step_kernel -s 8 -c 2 -o gpurun_out/r2_step_g4 python profiles/prof_step.py > gpurun_out/r2_step_g4.log 2>&1
  timeout 400 $NCU -k regex:reset_kernel -s 4 -c 1 -o gpurun_out/r2_reset python profiles/prof_step.py > gpurun_out/r2_reset.log 2>&1
  timeout 400 $NCU -k regex:env_step_kernel -s 8 -c 2 -o gpurun_out/r2_step_g2_n12 python profiles/prof_variants.py roundabout_2 12 8192 > gpurun_out/r2_step_g2_n12.log 2>&1
  timeout 400 $NCU -k regex:env_step_kernel -s 8 -c 2 -o gpurun_out/r2_step_ov1 python profiles/prof_variants.py cpm_entire 8 65536 is_obs_steering=1,is_observe_ref_path_other_agents=1 > gpurun_out/r2_step_ov1.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 3 --no-rollout --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
fi
if [ "$W" = san ] || [ "$W" = all ]; then
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/prof_variants.py roundabout_2 12 64 > gpurun_out/sanitizer_${tool}_n12.log 2>&1
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/prof_variants.py cpm_entire 8 64 is_ego_view=0,is_obs_steering=1 > gpurun_out/sanitizer_${tool}_ov1.log 2>&1
    tail -2 gpurun_out/sanitizer_${tool}_smoke.log gpurun_out/sanitizer_${tool}_n12.log gpurun_out/sanitizer_${tool}_ov1.log
  done
fi
ls -la gpurun_out
