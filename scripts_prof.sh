#!/bin/bash
mkdir -p gpurun_out
python profiles/profile_step.py > gpurun_out/prof_plain3.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wgrad4_tc_kernel|pn_lrelu" -s 20 -c 8 -o gpurun_out/wgrad_act_r1g python profiles/profile_step.py > gpurun_out/prof_ncu3.log 2>&1
tail -2 gpurun_out/prof_ncu3.log
