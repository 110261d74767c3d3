#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']//30)"
