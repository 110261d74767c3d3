#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== MINH=16"; PG_C4_MINH=16 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "test_conv_tc_matches_spec" 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
PG_C4_MINH=16 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python profiles/bench_conv.py --which fwd --iters 20 2>&1 | tail -13
