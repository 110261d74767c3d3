"""BASELINE config 2 on one B200: CorrectGenerator(512, 512) / CorrectDiscriminator(512), step 4
(32 px), batch 128, full Trainer iteration (D phase, Adam, G phase, Adam, EMA) — the 512-channel
3x3 convs on tcgen05 (N tiles of 256, stand-alone PixelNorm) against the CUDA-core kernels.

    python profiles/bench_wide.py [--batch 128] [--step 4] [--iters 5] [--simt-iters 2]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progan_b200  # noqa: E402


def run(wide, batch, step, iters, graph):
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc, K.wide_tc = "tc", True, wide
    K.invalidate_packs()
    torch.manual_seed(0)
    G = progan_b200.CorrectGenerator(512, 512, precision="bf16").cuda()
    D = progan_b200.CorrectDiscriminator(512, precision="bf16").cuda()
    Grun = progan_b200.CorrectGenerator(512, 512, precision="bf16").cuda()
    tr = progan_b200.Trainer(G, D, Grun, use_graph=graph)
    g = torch.Generator().manual_seed(1234)
    R = 2 * 2 ** step
    real = (torch.rand(batch, 3, R, R, generator=g) * 2 - 1).cuda()
    z = torch.randn(batch, 512, generator=g).cuda()
    eps = torch.rand(batch, 1, 1, 1, generator=g).cuda()
    for _ in range(3 if graph else 1):
        tr.step(real, z, eps, step, 0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tr.step(real, z, eps, step, 0.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    m = tr.read_metrics()
    return ms, m


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--step", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--simt-iters", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for an ncu launch list)")
    a = ap.parse_args()
    # 14 F_D + 3 F_G conv FLOPs per image (217.1 GFLOP/img at ch = 512, 32 px)
    gflop = {4: 217.1}.get(a.step)
    out = {"workload": "CorrectGenerator(512,512)/CorrectDiscriminator(512) step %d batch %d alpha 0.5"
                       % (a.step, a.batch)}
    ms, m = run(True, a.batch, a.step, a.iters, not a.no_graph)
    out["tcgen05_wide"] = {"ms_per_step": round(ms, 2), "img_per_s": round(a.batch / ms * 1e3, 1), "metrics": m}
    if gflop:
        out["tcgen05_wide"]["tflops"] = round(a.batch * gflop / ms, 1)
    if a.simt_iters > 0:
        ms2, m2 = run(False, a.batch, a.step, a.simt_iters, False)
        out["cuda_core"] = {"ms_per_step": round(ms2, 2), "img_per_s": round(a.batch / ms2 * 1e3, 1),
                            "metrics": m2}
    print(json.dumps(out))
