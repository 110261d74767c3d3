#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
