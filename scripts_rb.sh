#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_tc_matches or fused_avgpool or actbwd" 2>&1 | tail -2
timeout 300 python profiles/bench_conv.py --which fwd --iters 20 2>&1 | tail -13
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
