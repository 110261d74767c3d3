"""One launch each of the kernels whose ncu --set full summaries are committed under profiles/r2/
(after a warm-up launch): the FLOP-dominant conv4 layers, the weight-gradient kernel and the
HBM-bound kernels at their config-4 shapes (batch 64).
  python profiles/r2/capture_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200.kernels import ConvOp, EPI_PN_LRELU  # noqa: E402

K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
dev, bf, B = "cuda", torch.bfloat16, 64


def conv(res, cin, cout):
    x = torch.randn(B, res, res, cin, device=dev).to(bf)
    w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device=dev))
    b = torch.randn(cout, device=dev) * 0.1
    dy = torch.randn(B, res, res, cout, device=dev).to(bf)
    for _ in range(2):
        K.conv_fwd(x, w, b, ConvOp(3, 1), (2.0 / (cin * 9)) ** 0.5, EPI_PN_LRELU, 0.2)
    for _ in range(2):
        K.conv_wgrad(x, dy, tuple(w.shape), ConvOp(3, 1), 1.0)
    torch.cuda.synchronize()


conv(128, 64, 64)      # D block 1 conv.3 / its data gradient: the largest single layer
conv(64, 128, 128)     # D block 2 conv.3 (CTA-pair variant, resident half weights)
conv(128, 32, 32)      # G's last conv: the HBM-heavy end of the family
y = torch.randn(B, 128, 128, 64, device=dev).to(bf)
dy = torch.randn(B, 128, 128, 64, device=dev).to(bf)
dp = torch.randn(B, 64, 64, 64, device=dev).to(bf)
r = torch.rand(B, 128, 128, device=dev) + 0.5
img = torch.randn(B, 3, 128, 128, device=dev)
w32, b32 = torch.randn(32, 3, device=dev), torch.randn(32, device=dev)
a32 = torch.randn(B, 128, 128, 32, device=dev).to(bf)
for _ in range(2):
    K.pn_lrelu_bwd(dy, y, r, 0.2, True, False, True)
    K.pn_lrelu_bwd(dp, y, r, 0.2, True, True, True)
    K.pn_lrelu_bwd_bwd(dy, dy, y, r, 0.2, True, False)
    K.pw_expand(img, w32, b32, 32, 3, 1, 1.0, bf)
    K.pw_reduce(a32, torch.randn(3, 32, device=dev), None, 3, 1, 32, 1.0)
    K.pw_wgrad(a32, img, (32, 3), 3, 1, 1.0)
    K.upsample2(dp)
    K.upsample2_bwd(y)
torch.cuda.synchronize()
print("ok")
