"""Where does a conv4 launch spend its time?  Times each config-4 layer with the kernel's
experiment knobs (PG_DBG bits: 1 = no TMA stores, 2 = accumulator drain only, 8 = no activation
loads, 16 = no MMAs).  CUDA events, L2-rotating inputs.
  python profiles/r2/diag_conv4.py [--layers 128:32:32,...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["PG_DBG_LIVE"] = "1"     # the library re-reads PG_DBG on every launch
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200.kernels import ConvOp, EPI_PN_LRELU, EPI_LINEAR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--layers", default="128:32:32,128:32:64,128:64:32,128:64:64,64:64:128,64:128:64,64:128:128,32:128:128")
ap.add_argument("--dbg", default="0,1,2,8,16,18,24,26")
a = ap.parse_args()
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
dev = "cuda"
B = a.batch


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


dbgs = [int(v) for v in a.dbg.split(",")]
print("%-16s %8s %8s | " % ("layer", "hbm_us", "mma_us") + " ".join("dbg%-5d" % d for d in dbgs) + " | lin  pool")
for spec in a.layers.split(","):
    res, cin, cout = (int(v) for v in spec.split(":"))
    op = ConvOp(3, 1)
    w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device=dev))
    b = torch.randn(cout, device=dev) * 0.1
    nbytes = B * res * res * (cin + cout) * 2
    nsets = max(2, min(8, int(300e6 // nbytes) + 1))
    xs = [torch.randn(B, res, res, cin, device=dev).to(torch.bfloat16) for _ in range(nsets)]
    scale = (2.0 / (cin * 9)) ** 0.5
    flops = 2.0 * B * res * res * cin * cout * 9
    row = []
    for d in dbgs:
        os.environ["PG_DBG"] = str(d)
        row.append(timeit(lambda x: K.conv_fwd(x, w, b, op, scale, EPI_PN_LRELU, 0.2), [(x,) for x in xs]))
    os.environ["PG_DBG"] = "0"
    t_lin = timeit(lambda x: K.conv_fwd(x, w, None, op, scale, EPI_LINEAR, 0.2), [(x,) for x in xs])
    t_pool = timeit(lambda x: K.conv_fwd(x, w, b, op, scale, EPI_PN_LRELU, 0.2, pool_out=True), [(x,) for x in xs])
    print("%-16s %8.1f %8.1f | " % ("%d %d->%d" % (res, cin, cout), nbytes / 6544.3e3, flops / 1636.7e6)
          + " ".join("%-8.1f" % t for t in row) + " | %.1f %.1f" % (t_lin, t_pool), flush=True)
    del xs
