set -x
python profiles/kbench.py 65536 40 > gpurun_out/kb_base.log 2>&1
python profiles/kbench.py 8192 40 >> gpurun_out/kb_base.log 2>&1
KB_WRITE_OBS=1 python profiles/kbench.py 65536 40 >> gpurun_out/kb_base.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:env_step_kernel -s 8 -c 2 -o gpurun_out/base_r2 python profiles/prof_step.py > gpurun_out/ncu_base.log 2>&1
cat gpurun_out/kb_base.log
