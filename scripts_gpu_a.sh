#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/profile_calls.py > gpurun_out/calls_128.log 2>&1
python profiles/profile_step.py > gpurun_out/prof_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python profiles/profile_step.py > gpurun_out/prof_ncu.log 2>&1
python profiles/profile_step.py > gpurun_out/prof_plain2.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv4_tc_kernel -s 10 -c 4 -o gpurun_out/conv4_r1f python profiles/profile_step.py > gpurun_out/prof_ncu2.log 2>&1
tail -2 gpurun_out/prof_ncu2.log
