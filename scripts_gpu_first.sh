#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "not tc" > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad_tc" > gpurun_out/t2_wgrad_tc.log 2>&1; echo "wgrad_tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_tc" > gpurun_out/t3_conv_tc.log 2>&1; echo "conv_tc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q > gpurun_out/t4_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t*.log; do echo "== $f"; tail -n 4 $f; done
