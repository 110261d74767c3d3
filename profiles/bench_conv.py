"""Per-layer microbenchmark of the tcgen05 conv / wgrad kernels (CUDA events, L2 flushed
between iterations by cycling through buffers larger than L2).
  python profiles/bench_conv.py [--batch 64] [--which fwd,wgrad]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200.kernels import ConvOp, EPI_PN_LRELU, EPI_LINEAR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--which", default="fwd,wgrad")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--layers", default="", help="comma list like 128:64:64 to restrict")
a = ap.parse_args()
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
dev = "cuda"
B = a.batch
# (res, Cin, Cout) of config 4 at 128px: D blocks + G blocks (+ their dgrad forms = swapped)
LAYERS = [(128, 32, 64), (128, 64, 64), (128, 64, 32), (128, 32, 32), (64, 64, 128), (64, 128, 128),
          (64, 128, 64), (64, 64, 64), (32, 128, 128), (16, 128, 128), (8, 128, 128), (4, 128, 128)]
PEAK = 1384.2


def timeit(fn, sets):
    for s in sets[:2]:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3   # us


print("%-22s %10s %10s %8s" % ("layer", "us", "TFLOP/s", "frac"))
if a.layers:
    LAYERS = [tuple(int(v) for v in l.split(":")) for l in a.layers.split(",")]
for res, cin, cout in LAYERS:
    op = ConvOp(3, 1)
    w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device=dev))
    b = torch.randn(cout, device=dev) * 0.1
    nbytes = B * res * res * (cin + cout) * 2
    nsets = max(2, min(8, int(300e6 // nbytes) + 1))     # > 126 MB L2 in rotation
    xs = [torch.randn(B, res, res, cin, device=dev).to(torch.bfloat16) for _ in range(nsets)]
    dys = [torch.randn(B, res, res, cout, device=dev).to(torch.bfloat16) for _ in range(nsets)]
    flops = 2.0 * B * res * res * cin * cout * 9
    scale = (2.0 / (cin * 9)) ** 0.5
    if "fwd" in a.which:
        t = timeit(lambda x: K.conv_fwd(x, w, b, op, scale, EPI_PN_LRELU, 0.2), [(x,) for x in xs])
        print("%-22s %10.1f %10.1f %8.3f" % ("fwd %d %d->%d" % (res, cin, cout), t, flops / t / 1e6, flops / t / 1e6 / PEAK))
    if "wgrad" in a.which:
        t = timeit(lambda x, dy: K.conv_wgrad(x, dy, tuple(w.shape), op, scale), list(zip(xs, dys)))
        print("%-22s %10.1f %10.1f %8.3f" % ("wgrad %d %d->%d" % (res, cin, cout), t, flops / t / 1e6, flops / t / 1e6 / PEAK))
    del xs, dys
