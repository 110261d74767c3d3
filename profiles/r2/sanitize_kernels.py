"""Every kernel family once at small shapes, for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python profiles/r2/sanitize_kernels.py
One full Trainer iteration at 32 px (batch 4, 32 channels: conv4 resident / pooled / CTA-pair /
fused-backward variants, the small-map conv_tc / wgrad_tc kernels, weight-gradient kernels with
deferred unpack, PixelNorm first and second order, 1x1 heads, resampling, minibatch-stddev,
gradient penalty, multi-tensor Adam, EMA) plus the 128-channel variants on their own."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200.kernels import ConvOp, EPI_PN_LRELU  # noqa: E402

dev, bf = "cuda", torch.bfloat16
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
torch.manual_seed(0)
# (1) whole iteration, small model: 32 channels, step 3 (32 px), fade active
G = progan_b200.Generator(32, 32, tanh=False).to(dev)
D = progan_b200.Discriminator(32).to(dev)
R = progan_b200.Generator(32, 32, tanh=False).to(dev)
tr = progan_b200.Trainer(G, D, R, use_graph=False)
g = torch.Generator().manual_seed(1)
real = (torch.rand(4, 3, 32, 32, generator=g) * 2 - 1).to(dev)
z = torch.randn(4, 32, generator=g).to(dev)
eps = torch.rand(4, 1, 1, 1, generator=g).to(dev)
for fuse in (0, 128):
    K.fuse_actbwd_max_cout = fuse
    tr.step(real, z, eps, 3, 0.5)
torch.cuda.synchronize()
print("iteration ok", tr.read_metrics())
# (2) the wide variants of the tcgen05 kernels on their own
for (n, res, cin, cout) in [(2, 32, 128, 128), (2, 16, 128, 128), (2, 32, 128, 64), (2, 32, 64, 128), (2, 64, 64, 64),
                            (1, 64, 32, 32), (4, 8, 128, 128), (4, 4, 128, 128)]:
    x = torch.randn(n, res, res, cin, device=dev).to(bf)
    w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device=dev))
    b = torch.randn(cout, device=dev) * 0.1
    dy = torch.randn(n, res, res, cout, device=dev).to(bf)
    y, r, yp = K.conv_fwd(x, w, b, ConvOp(3, 1), 0.05, EPI_PN_LRELU, 0.2, pool_out=True)
    K.conv_wgrad(x, dy, tuple(w.shape), ConvOp(3, 1), 1.0)
    K.fuse_actbwd_max_cout = 128
    K.conv_dgrad_actbwd(dy, w, ConvOp(3, 1).adjoint(), 0.05, torch.randn(n, res, res, cin, device=dev).to(bf),
                        torch.rand(n, res, res, device=dev) + 0.5, 0.2, True, torch.zeros(cin, device=dev))
    K.fuse_actbwd_max_cout = 0
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
