timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_now.json
python -c "
import json; d=json.load(open('gpurun_out/bench_now.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline'])"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('eager', {k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})"
for r in 64 32 16 8; do timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --res $r 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])"; done
