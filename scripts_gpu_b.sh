#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
