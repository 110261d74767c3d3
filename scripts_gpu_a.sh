#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_now.json
timeout 600 python profiles/profile_calls.py > gpurun_out/calls_128.log 2>&1
python profiles/profile_step.py > gpurun_out/prof_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python profiles/profile_step.py > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gpu.log; cat gpurun_out/bench_now.json | cut -c1-400
