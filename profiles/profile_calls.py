"""Per-call warm timing of one eager training iteration: every C-ABI call is bracketed by CUDA
events (a measurement harness around the product path, not part of it).
  python profiles/profile_calls.py [--res 128] [--batch 64] [--reps 3]
Prints time per (entry point, integer arguments) so the slow layers can be named.
"""
import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import progan_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--alpha", type=float, default=0.5)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--correct", type=int, default=0, help="channel width of CorrectGenerator/Discriminator "
                "(step = log2(res) - 1) instead of train.py's models")
a = ap.parse_args()
dev = torch.device("cuda", 0)
step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[a.res]
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
torch.manual_seed(0)
zdim = 128
if a.correct:
    zdim, step = a.correct, step + 1
    G = progan_b200.CorrectGenerator(zdim, a.correct).to(dev)
    D = progan_b200.CorrectDiscriminator(a.correct).to(dev)
    Gr = progan_b200.CorrectGenerator(zdim, a.correct).to(dev)
else:
    G = progan_b200.Generator(128, 128, tanh=False).to(dev)
    D = progan_b200.Discriminator(128).to(dev)
    Gr = progan_b200.Generator(128, 128, tanh=False).to(dev)
tr = progan_b200.Trainer(G, D, Gr, use_graph=False)
g = torch.Generator().manual_seed(1234)
real = (torch.rand(a.batch, 3, a.res, a.res, generator=g) * 2 - 1).to(dev)
z = torch.randn(a.batch, zdim, generator=g).to(dev)
eps = torch.rand(a.batch, 1, 1, 1, generator=g).to(dev)
for _ in range(2):
    tr.step(real, z, eps, step, a.alpha)
torch.cuda.synchronize()

records = []
orig_call = K._call


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *args)
    e1.record()
    key = tuple(x for x in args if isinstance(x, int) and not isinstance(x, bool) and x < 1 << 20)
    records.append((name, key, e0, e1))


K._call = timed_call
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(a.reps):
    tr.step(real, z, eps, step, a.alpha)
t1.record()
torch.cuda.synchronize()
K._call = orig_call
agg = defaultdict(lambda: [0, 0.0])
byname = defaultdict(lambda: [0, 0.0])
for name, key, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    agg[(name, key)][0] += 1
    agg[(name, key)][1] += ms
    byname[name][0] += 1
    byname[name][1] += ms
tot = sum(v[1] for v in agg.values()) / a.reps
print("eager iteration %.3f ms; sum of bracketed calls %.3f ms (per iteration, %d reps)"
      % (t0.elapsed_time(t1) / a.reps, tot, a.reps))
print("---- by entry point")
for n, (c, ms) in sorted(byname.items(), key=lambda t: -t[1][1]):
    print("%-24s %5d calls %9.1f us %5.1f%%" % (n, c // a.reps, 1e3 * ms / a.reps, 100 * ms / a.reps / tot))
print("---- by entry point and integer arguments")
for (n, key), (c, ms) in sorted(agg.items(), key=lambda t: -t[1][1])[:70]:
    print("%-22s x%-3d %8.1f us  (%6.1f each)  %s" % (n, c // a.reps, 1e3 * ms / a.reps, 1e3 * ms / c, key))
