"""Microbenchmark of the bandwidth kernels at the 128px shapes (CUDA events, rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import progan_b200
K = progan_b200.get_kernels()
dev = "cuda"
def timeit(fn, n=10):
    fn(0); fn(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (N, H, W, C) in [(64, 128, 128, 64), (64, 128, 128, 32), (64, 64, 64, 128)]:
    ys = [torch.randn(N, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)]
    dys = [torch.randn(N, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)]
    dps = [torch.randn(N, H // 2, W // 2, C, device=dev).to(torch.bfloat16) for _ in range(3)]
    ts = [torch.randn(N, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)]
    r = torch.rand(N, H, W, device=dev) + 0.5
    nel = N * H * W * C
    t = timeit(lambda i: K.pn_lrelu_bwd(dys[i % 3], ys[i % 3], r, 0.2, True, False, True))
    print("pn_bwd        %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, (6 * nel + 4 * nel / C) / t / 1e3))
    t = timeit(lambda i: K.pn_lrelu_bwd(dys[i % 3], ys[i % 3], r, 0.2, True, False, False))
    print("pn_bwd nocs   %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, (6 * nel + 4 * nel / C) / t / 1e3))
    t = timeit(lambda i: K.pn_lrelu_bwd(dps[i % 3], ys[i % 3], r, 0.2, True, True, True))
    print("pn_bwd pooled %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, (4.5 * nel + 4 * nel / C) / t / 1e3))
    t = timeit(lambda i: K.pn_lrelu_bwd_bwd(ts[i % 3], dys[i % 3], ys[i % 3], r, 0.2, True, False))
    print("pn_bwd_bwd    %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, (10 * nel + 4 * nel / C) / t / 1e3))
    t = timeit(lambda i: K.avgpool2(ys[i % 3]))
    print("avgpool2      %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, 2.5 * nel / t / 1e3))
    t = timeit(lambda i: K.colsum(ys[i % 3]))
    print("colsum        %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, 2 * nel / t / 1e3))
    t = timeit(lambda i: K.upsample2(dps[i % 3]))
    print("upsample2     %s %7.1f us  %.0f GB/s" % ((N, H, W, C), t, 2.5 * nel / t / 1e3))
img = [torch.randn(64, 3, 128, 128, device=dev) for _ in range(3)]
w = torch.randn(32, 3, device=dev); b = torch.randn(32, device=dev)
t = timeit(lambda i: K.pw_expand(img[i % 3], w, b, 32, 3, 1, 1.0, torch.bfloat16))
print("pw_expand 3->32 @128  %7.1f us  %.0f GB/s" % (t, (64 * 128 * 128 * (32 * 2 + 12)) / t / 1e3))
act = [torch.randn(64, 128, 128, 32, device=dev).to(torch.bfloat16) for _ in range(3)]
t = timeit(lambda i: K.pw_wgrad(act[i % 3], img[i % 3], (32, 3), 3, 1, 1.0))
print("pw_wgrad 32x3 @128    %7.1f us  %.0f GB/s" % (t, (64 * 128 * 128 * (32 * 2 + 12)) / t / 1e3))
wt = torch.randn(3, 32, device=dev); bt = torch.randn(3, device=dev)
t = timeit(lambda i: K.pw_reduce(act[i % 3], wt, bt, 3, 1, 32, 1.0))
print("pw_reduce 32->3 @128  %7.1f us  %.0f GB/s" % (t, (64 * 128 * 128 * (32 * 2 + 12)) / t / 1e3))
