"""BASELINE config 5 shape on one B200: ConditionalCorrectGenerator(512, 14 classes, 512, max_step=6) /
ConditionalCorrectDiscriminatorWgangp(512, 14), step 6 (128 px), full Trainer iteration with labels.

    python profiles/bench_cond.py [--batch 16] [--step 6] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progan_b200  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--step", type=int, default=6)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--ch", type=int, default=512)
    a = ap.parse_args()
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "tc", True
    torch.manual_seed(0)
    mk = lambda: progan_b200.ConditionalCorrectGenerator(a.ch, 14, a.ch, max_step=6, precision="bf16").cuda()
    G, Grun = mk(), mk()
    D = progan_b200.ConditionalCorrectDiscriminatorWgangp(a.ch, 14, precision="bf16").cuda()
    tr = progan_b200.Trainer(G, D, Grun, use_graph=True)
    g = torch.Generator().manual_seed(1234)
    R = 2 * 2 ** a.step
    real = (torch.rand(a.batch, 3, R, R, generator=g) * 2 - 1).cuda()
    z = torch.randn(a.batch, a.ch, generator=g).cuda()
    eps = torch.rand(a.batch, 1, 1, 1, generator=g).cuda()
    label = torch.randint(0, 14, (a.batch,), generator=g).cuda()
    for _ in range(3):
        tr.step(real, z, eps, a.step, 0.5, label=label)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        tr.step(real, z, eps, a.step, 0.5, label=label)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    out = {"workload": "ConditionalCorrectGenerator(%d,14,%d,max_step=6)/ConditionalCorrectDiscriminatorWgangp(%d,14) "
                       "step %d (%d px) batch %d alpha 0.5" % (a.ch, a.ch, a.ch, a.step, R, a.batch),
           "ms_per_step": round(ms, 2), "img_per_s": round(a.batch / ms * 1e3, 1), "metrics": tr.read_metrics(),
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if a.ch == 512 and a.step == 6:
        out["tflops"] = round(a.batch * 1116.0 / ms, 1)     # 1116 GFLOP/img
    print(json.dumps(out))
