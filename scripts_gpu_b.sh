#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph 2>&1 | tail -1 | cut -c1-200
