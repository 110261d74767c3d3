#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train.py -q -k "bf16 or fast_paths" 2>&1 | tail -1; done
