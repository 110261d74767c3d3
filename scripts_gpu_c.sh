#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train.py -q -k "bf16 or fast_paths" 2>&1 | grep -E "^E   .*(Assert|assert)|passed|failed|^FAILED" | cut -c1-700; done
timeout 300 python profiles/dbg_determinism.py 2>&1 | grep -v "^    " | tail -9
