#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q > gpurun_out/t4_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/t5_train.log 2>&1; echo "train rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench_graph rc=$?" >> gpurun_out/summary.txt
python profiles/profile_step.py > gpurun_out/prof_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python profiles/profile_step.py > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/summary.txt
for f in gpurun_out/t*.log gpurun_out/bench*.log; do echo "== $f"; tail -n 3 $f | cut -c1-330; done
