import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, progan_b200, helpers
from progan_b200.kernels import ConvOp, EPI_LINEAR
K = progan_b200.get_kernels(); K.conv_impl = "tc"; K.fuse_actbwd_min_cout = 32
dev = "cuda"
torch.manual_seed(0)
for (N, H, W, Cin, Cout) in [(64, 128, 128, 64, 64), (64, 64, 64, 128, 128), (64, 32, 32, 128, 128), (64, 16, 16, 128, 128), (128, 128, 128, 64, 64), (64, 128, 128, 32, 32), (64, 64, 64, 64, 64)]:
    op = ConvOp(3, 1, True, True)
    x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
    w = torch.nn.Parameter(torch.randn(Cin, Cout, 3, 3, device=dev))
    a = torch.randn(N, H, W, Cout, device=dev)
    r = torch.rsqrt((a ** 2).mean(-1) + 1e-8)
    pn = a * r.unsqueeze(-1)
    y = torch.where(pn > 0, pn, 0.2 * pn).to(torch.bfloat16)
    outs = []
    for rep in range(4):
        cs = torch.zeros(Cout, device=dev)
        da = K.conv_dgrad_actbwd(x, w, op, 0.05, y, r, 0.2, True, cs)
        torch.cuda.synchronize()
        outs.append((da.clone(), cs.clone()))
    dh, _ = K.conv_fwd(x, w, None, op, 0.05, EPI_LINEAR)
    cs_ref = torch.zeros(Cout, device=dev)
    da_ref, _ = K.pn_lrelu_bwd(dh, y, r, 0.2, True, False, False, cs_ref)
    same = all(torch.equal(outs[0][0], o[0]) for o in outs[1:])
    csd = max(helpers.rel(o[1], outs[0][1]) for o in outs[1:])
    # timing
    def t(fn, n=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    cs = torch.zeros(Cout, device=dev)
    tf = t(lambda: K.conv_dgrad_actbwd(x, w, op, 0.05, y, r, 0.2, True, cs))
    tn = t(lambda: K.conv_dgrad_actbwd(x, w, op, 0.05, y, r, 0.2, True, None))
    t1 = t(lambda: K.conv_fwd(x, w, None, op, 0.05, EPI_LINEAR))
    t2 = t(lambda: K.pn_lrelu_bwd(dh, y, r, 0.2, True, False, False, cs_ref))
    cs_ref = outs[0][1]
    print((N, H, W, Cin, Cout), "bitwise same:", same, "colsum run-to-run %.1e" % csd,
          "vs 2-kernel: da %.2e cs %.2e" % (helpers.rel(outs[0][0], da_ref), helpers.rel(outs[0][1], cs_ref)),
          "| fused %.1f us (no colsum %.1f) vs conv %.1f + act %.1f" % (tf, tn, t1, t2))
