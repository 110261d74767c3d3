timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/t_k.log 2>&1; tail -2 gpurun_out/t_k.log
python profiles/bench_conv.py 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
