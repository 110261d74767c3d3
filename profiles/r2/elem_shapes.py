"""The HBM-bound kernels of one training iteration at their config-4 shapes (batch 64, 128 px),
each launched `--reps` times on rotating buffers: for ncu captures and CUDA-event timing.
  python profiles/r2/elem_shapes.py [--reps 3] [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
K = progan_b200.get_kernels()
dev = "cuda"
B = 64
bf = torch.bfloat16
PEAK = 6544.3


def rot(shape, dtype=bf, n=3):
    return [torch.randn(shape, device=dev).to(dtype) for _ in range(n)]


cases = []


def case(name, nbytes, fn):
    cases.append((name, nbytes, fn))


img128 = rot((B, 3, 128, 128), torch.float32)
img64 = rot((B, 3, 64, 64), torch.float32)
act128_32 = rot((B, 128, 128, 32))
act128_64 = rot((B, 128, 128, 64))
act64_64 = rot((B, 64, 64, 64))
act64_128 = rot((B, 64, 64, 128))
dp64_64 = rot((B, 64, 64, 64))
r128 = torch.rand(B, 128, 128, device=dev) + 0.5
w32 = torch.randn(32, 3, device=dev)
b32 = torch.randn(32, device=dev)
w64 = torch.randn(64, 3, device=dev)
b64 = torch.randn(64, device=dev)
wt32 = torch.randn(3, 32, device=dev)
bt3 = torch.randn(3, device=dev)
px = B * 128 * 128
case("pw_expand 3->32 @128", px * (32 * 2 + 12), lambda i: K.pw_expand(img128[i % 3], w32, b32, 32, 3, 1, 1.0, bf))
case("pw_expand 3->64 @64", px // 4 * (64 * 2 + 12), lambda i: K.pw_expand(img64[i % 3], w64, b64, 64, 3, 1, 1.0, bf))
case("pw_reduce 32->3 @128", px * (32 * 2 + 12), lambda i: K.pw_reduce(act128_32[i % 3], wt32, bt3, 3, 1, 32, 1.0))
case("pw_reduce 64->3 @64", px // 4 * (64 * 2 + 12), lambda i: K.pw_reduce(act64_64[i % 3], torch.randn(3, 64, device=dev), bt3, 3, 1, 64, 1.0))
case("pw_wgrad 32x3 @128", px * (32 * 2 + 12), lambda i: K.pw_wgrad(act128_32[i % 3], img128[i % 3], (32, 3), 3, 1, 1.0))
case("upsample2 64ch 64->128", px * 64 * 2 * 1.25, lambda i: K.upsample2(act64_64[i % 3]))
case("upsample2_bwd 64ch 128->64", px * 64 * 2 * 1.25, lambda i: K.upsample2_bwd(act128_64[i % 3]))
case("upsample2 nchw fp32 img 64->128", px * 3 * 4 * 1.25, lambda i: K.upsample2(img64[i % 3], "nchw"))
case("avgpool2 nchw fp32 img 128->64", px * 3 * 4 * 1.25, lambda i: K.avgpool2(img128[i % 3], "nchw"))
case("pn_lrelu_bwd 64ch @128", px * 64 * 2 * 3, lambda i: K.pn_lrelu_bwd(act128_64[i % 3], act128_64[(i + 1) % 3], r128, 0.2, True, False, True))
case("pn_lrelu_bwd pooled 64ch @128", px * 64 * 2 * 2.25, lambda i: K.pn_lrelu_bwd(dp64_64[i % 3], act128_64[(i + 1) % 3], r128, 0.2, True, True, True))
case("pn_lrelu_bwd_bwd 64ch @128", px * 64 * 2 * 5, lambda i: K.pn_lrelu_bwd_bwd(act128_64[i % 3], act128_64[(i + 1) % 3], act128_64[(i + 2) % 3], r128, 0.2, True, False))
alpha = torch.full((), 0.5, device=dev)
case("blend 64ch @64", px // 4 * 64 * 2 * 3, lambda i: K.blend(act64_64[i % 3], act64_64[(i + 1) % 3], alpha))
# optimiser-side kernels on D-sized buckets (1.5 M parameters)
n_par = 1536 * 1024
p_, g_, v_ = (torch.randn(n_par, device=dev) for _ in range(3))
v_.abs_()
chunks = torch.tensor([(s, min(2048, n_par - s), 0, 0) for s in range(0, n_par, 2048)], dtype=torch.int32, device=dev)
steps = torch.ones(1, device=dev)
case("adam_multi 1.5M params", n_par * 4 * 5, lambda i: K.adam_multi(p_, g_, None, v_, chunks, steps, 1e-3, 0.0, 0.99, 1e-8))
# wgrad unpack: twelve 128x128x9 workspaces (the D bucket's trunk)
ws = [torch.randn(9 * 128 * 128, device=dev) for _ in range(12)]
dws = [torch.zeros(128, 128, 3, 3, device=dev) for _ in range(12)]
ents = [_lib.UnpackEntry(w_.data_ptr(), d_.data_ptr(), 128, 128, 128, 128, 9, 0, 0, 0, 0.1, 0.0) for w_, d_ in zip(ws, dws)]
tab = K._upload(ents, _lib.UnpackEntry, torch.device(dev))
case("wgrad_unpack_multi 12 x 128x128x9", 12 * 9 * 128 * 128 * 4 * 4, lambda i: K._call("pg_wgrad_unpack_multi", tab.data_ptr(), 12, K._stream()))
pars = [torch.nn.Parameter(torch.randn(128, 128, 3, 3, device=dev)) for _ in range(12)]
from progan_b200.kernels import ConvOp, WL_CO_TAP_CI  # noqa: E402
for p in pars:
    K.packed(p, ConvOp(3, 1), WL_CO_TAP_CI, bf)
    K.packed(p, ConvOp(3, 1).adjoint(), WL_CO_TAP_CI, bf)
case("pack_weight_multi 24 x 128x128x9", 24 * 9 * 128 * 128 * 6, lambda i: K.refresh_packs(pars))

for name, nbytes, fn in cases:
    fn(0)
    torch.cuda.synchronize()
    if a.time:
        # ten launches on rotating buffers replayed as ONE CUDA graph: no host launch cost in the
        # timed region (an eager loop of 20-30 us kernels measures the Python/ctypes call rate)
        fn(1); fn(2)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for i in range(10):
                fn(i)
        gph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print("%-38s %8.1f us %8.0f GB/s  %.2f of %.0f" % (name, us, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK, PEAK))
    else:
        for i in range(a.reps):
            fn(i)
        torch.cuda.synchronize()
print("done")
