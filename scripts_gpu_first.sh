#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm_and_padded or mbstd" > gpurun_out/t0_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q > gpurun_out/t4_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/t5_train.log 2>&1; echo "train rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench_graph rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t*.log gpurun_out/bench*.log; do echo "== $f"; tail -n 4 $f | cut -c1-400; done
