#!/bin/bash
timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -1
