#!/bin/bash
# bash profiles/r2/scale_run.sh N  — bench.py at every progressive resolution on N GPUs of one box
# (weak scaling, 64 images per GPU), then the data-parallel check at 128 px; lines go to
# gpurun_out/scale_steps_n$N.jsonl and gpurun_out/ddp_check_${N}gpu.txt.  An N-GPU call is charged N x its
# wall time: give gpurun a --timeout that covers only what you can afford (this script's own
# worst case is 5 x 120 s + 150 s).
N=$1
OUT=gpurun_out/scale_steps_n$N.jsonl
rm -f $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for r in 128 64 32 16 8; do
  timeout 120 $RUN bench.py --gpus $N --res $r --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 >> $OUT
done
timeout 150 $RUN profiles/r2/ddp_check.py --steps 30 2>&1 | grep -E "ms/iter|rel diff|DDP_CHECK" > gpurun_out/ddp_check_${N}gpu.txt
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(d["metric"], "n=%d" % d["n_gpus"], d["ms_per_step"], d["value"], d.get("replicas_identical"))
PY
cat gpurun_out/ddp_check_${N}gpu.txt
