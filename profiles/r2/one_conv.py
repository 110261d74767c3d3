"""One conv4 launch (plus a warm-up) for ncu: python profiles/r2/one_conv.py RES CIN COUT [BATCH]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import progan_b200  # noqa: E402
from progan_b200.kernels import ConvOp, EPI_PN_LRELU  # noqa: E402

res, cin, cout = (int(v) for v in sys.argv[1:4])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
K = progan_b200.get_kernels()
K.conv_impl = "tc"
x = torch.randn(B, res, res, cin, device="cuda").to(torch.bfloat16)
w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device="cuda"))
b = torch.randn(cout, device="cuda") * 0.1
for _ in range(3):
    y, r = K.conv_fwd(x, w, b, ConvOp(3, 1), (2.0 / (cin * 9)) ** 0.5, EPI_PN_LRELU, 0.2)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
