"""Kernel timeline of graph-replayed training iterations (CUPTI through torch.profiler): per-kernel
totals, the span of one iteration, how much of it has 0 / 1 / >=2 kernels in flight, and the
time each stream is busy.  Writes gpurun_out/trace_step.json (chrome trace) when asked.
  python profiles/r2/trace_step.py [--res 128] [--batch 64] [--iters 3] [--dump]"""
import argparse
import json
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import progan_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--alpha", type=float, default=0.5)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--dump", action="store_true")
ap.add_argument("--serial", action="store_true", help="no side streams (overlap_wgrad / overlap_passes off)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[a.res]
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
torch.manual_seed(0)
G = progan_b200.Generator(128, 128, tanh=False).to(dev)
D = progan_b200.Discriminator(128).to(dev)
Gr = progan_b200.Generator(128, 128, tanh=False).to(dev)
tr = progan_b200.Trainer(G, D, Gr, use_graph=not a.no_graph, overlap_wgrad=not a.serial,
                         overlap_passes=not a.serial)
g = torch.Generator().manual_seed(1234)
real = (torch.rand(a.batch, 3, a.res, a.res, generator=g) * 2 - 1).to(dev)
z = torch.randn(a.batch, 128, generator=g).to(dev)
eps = torch.rand(a.batch, 1, 1, 1, generator=g).to(dev)
for _ in range(4):
    tr.step(real, z, eps, step, a.alpha)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.step(real, z, eps, step, a.alpha)
e1.record()
torch.cuda.synchronize()
print("unprofiled: %.3f ms/iteration" % (e0.elapsed_time(e1) / 10))
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.iters):
        tr.step(real, z, eps, step, a.alpha)
    torch.cuda.synchronize()
path = "/tmp/trace_step.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and e.get("ph") == "X"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
span = max(e["ts"] + e["dur"] for e in ev) - t0
print("kernels=%d span=%.1f us per iteration=%.1f us  sum(dur)=%.1f us per iteration"
      % (len(ev), span, span / a.iters, sum(e["dur"] for e in ev) / a.iters))
agg = defaultdict(lambda: [0, 0.0])
for e in ev:
    n = re.sub(r"<.*", "", re.sub(r"\(.*", "", e["name"]))
    agg[n][0] += 1
    agg[n][1] += e["dur"]
print("%-44s %6s %10s" % ("kernel", "n/iter", "us/iter"))
for n, (c, d) in sorted(agg.items(), key=lambda t: -t[1][1])[:28]:
    print("%-44s %6.1f %10.1f" % (n[:44], c / a.iters, d / a.iters))
# concurrency histogram over the span
pts = []
for e in ev:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
hist = defaultdict(float)
cur, last = 0, pts[0][0]
for t, d in pts:
    hist[min(cur, 3)] += t - last
    cur += d
    last = t
print("in flight: " + "  ".join("%s: %.1f us/iter" % (("%d" % k) if k < 3 else ">=3", v / a.iters)
                                 for k, v in sorted(hist.items())))
by_stream = defaultdict(float)
for e in ev:
    by_stream[e["args"].get("stream", e.get("tid"))] += e["dur"]
print("busy per stream (us/iter): " + "  ".join("%s: %.0f" % (k, v / a.iters) for k, v in sorted(by_stream.items(), key=lambda t: -t[1])))
if a.dump:
    os.makedirs("gpurun_out", exist_ok=True)
    slim = [dict(name=re.sub(r"\(.*", "", e["name"])[:60], ts=round(e["ts"] - t0, 2), dur=round(e["dur"], 2),
                 stream=e["args"].get("stream"), grid=e["args"].get("grid"), block=e["args"].get("block"))
            for e in ev]
    json.dump(slim, open("gpurun_out/trace_step.json", "w"))
