timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad_tc" > gpurun_out/t_k.log 2>&1; tail -3 gpurun_out/t_k.log
python profiles/bench_conv.py --which wgrad 2>&1
