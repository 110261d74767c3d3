#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/profile_calls.py > gpurun_out/calls_128.log 2>&1
python profiles/profile_step.py > gpurun_out/prof_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python profiles/profile_step.py > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_plain.log
