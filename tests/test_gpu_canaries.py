"""GPU: out-of-bounds WRITE check with canaries.  compute-sanitizer is closed on this GPU pool
(gpurun: "closed on this pool and stays closed"), so every output buffer of the main kernel
families is carved out of the middle of a sentinel-filled allocation through the C-ABI directly
(caller-owned buffers), and the sentinels on both sides must survive the launch — at shapes that
are not multiples of the kernels' tile sizes as well."""
import ctypes

import pytest
import torch

import progan_b200
from progan_b200 import _lib
from progan_b200.kernels import ConvOp, WL_CO_TAP_CI

pytestmark = pytest.mark.gpu
DEV = "cuda"
PAD = 4096                      # sentinel elements on each side
SENT = 12345.0


class Guarded:
    """A tensor of `shape` in the middle of a sentinel-filled buffer."""

    def __init__(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        self.buf = torch.full((n + 2 * PAD,), SENT, device=DEV, dtype=dtype)
        self.t = self.buf[PAD:PAD + n].view(shape)
        assert self.t.data_ptr() % 16 == 0

    def check(self, what):
        assert bool((self.buf[:PAD] == SENT).all()) and bool((self.buf[-PAD:] == SENT).all()), \
            "%s wrote outside its output buffer" % what


def _stream():
    return torch.cuda.current_stream().cuda_stream


CONVS = [(1, 16, 16, 32, 32), (3, 16, 16, 64, 64), (2, 32, 32, 128, 128), (2, 32, 16, 128, 64), (1, 64, 64, 32, 64),
         (5, 8, 8, 128, 128), (3, 4, 4, 160, 128), (2, 16, 16, 64, 128)]


@pytest.mark.parametrize("N,H,W,Cin,Cout", CONVS)
def test_conv_and_weight_gradient_stay_inside_their_buffers(N, H, W, Cin, Cout):
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "tc", True
    lib = K.lib
    bf = torch.bfloat16
    x = torch.randn(N, H, W, Cin, device=DEV).to(bf)
    w = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device=DEV))
    bias = torch.randn(Cout, device=DEV)
    wp = K.packed(w, ConvOp(3, 1), WL_CO_TAP_CI, bf)
    y, r = Guarded((N, H, W, Cout), bf), Guarded((N, H, W), torch.float32)
    pool_ok = H % 16 == 0 and W % 8 == 0 and Cout in (32, 64, 128)
    yp = Guarded((N, H // 2, W // 2, Cout), bf) if pool_ok else None
    rc = lib.pg_conv_tc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.t.data_ptr(), r.t.data_ptr(), N, H, W, Cin,
                        Cout, Cout, 9, Cout, ctypes.c_float(0.05), 1, ctypes.c_float(0.2),
                        yp.t.data_ptr() if yp else None, _stream())
    _lib.check(rc, "pg_conv_tc")
    dy = torch.randn(N, H, W, Cout, device=DEV).to(bf)
    dw, ws = Guarded((Cout, Cin, 3, 3), torch.float32), Guarded((9 * Cin * Cout,), torch.float32)
    rc = lib.pg_conv_wgrad_tc(x.data_ptr(), dy.data_ptr(), dw.t.data_ptr(), ws.t.data_ptr(), N, H, W, Cin, Cout,
                              Cin, Cout, 9, 0, ctypes.c_float(1.0), 0, 0, 0, _stream())
    _lib.check(rc, "pg_conv_wgrad_tc")
    if pool_ok and Cout <= 128 and H >= 16:          # fused activation-backward epilogue
        yprev = torch.randn(N, H, W, Cin, device=DEV).to(bf)
        rprev = torch.rand(N, H, W, device=DEV) + 0.5
        wpa = K.packed(w, ConvOp(3, 1).adjoint(), WL_CO_TAP_CI, bf)
        da, cs = Guarded((N, H, W, Cin), bf), Guarded((Cin,), torch.float32)
        cs.t.zero_()
        rc = lib.pg_conv_tc_actbwd(dy.data_ptr(), wpa.data_ptr(), da.t.data_ptr(), N, H, W, Cout, Cin,
                                   ctypes.c_float(0.05), yprev.data_ptr(), rprev.data_ptr(), ctypes.c_float(0.2), 1,
                                   cs.t.data_ptr(), _stream())
        if rc == 0:
            torch.cuda.synchronize()
            da.check("pg_conv_tc_actbwd(da)")
            cs.check("pg_conv_tc_actbwd(colsum)")
    torch.cuda.synchronize()
    for g_, what in ((y, "pg_conv_tc(y)"), (r, "pg_conv_tc(r)"), (dw, "pg_conv_wgrad_tc(dw)"), (ws, "pg_conv_wgrad_tc(ws)")):
        g_.check(what)
    if yp:
        yp.check("pg_conv_tc(y_pool)")


@pytest.mark.parametrize("N,H,W,C", [(3, 10, 6, 32), (2, 16, 16, 64), (1, 4, 4, 128), (5, 7, 9, 8)])
def test_elementwise_kernels_stay_inside_their_buffers(N, H, W, C):
    K = progan_b200.get_kernels()
    lib = K.lib
    bf = torch.bfloat16
    st = _stream()
    act = torch.randn(N, H, W, C, device=DEV).to(bf)
    img = torch.randn(N, 3, H, W, device=DEV)
    w, b = torch.randn(C, 3, device=DEV), torch.randn(C, device=DEV)
    out = Guarded((N, H, W, C), bf)
    _lib.check(lib.pg_pw_expand(img.data_ptr(), w.data_ptr(), b.data_ptr(), out.t.data_ptr(), N, H * W, 3, C, 3, 1,
                                ctypes.c_float(1.0), 1, st), "pg_pw_expand")
    oimg = Guarded((N, 3, H, W), torch.float32)
    wt = torch.randn(3, C, device=DEV)
    _lib.check(lib.pg_pw_reduce(act.data_ptr(), wt.data_ptr(), None, oimg.t.data_ptr(), N, H * W, 3, C, 1, C,
                                ctypes.c_float(1.0), 1, st), "pg_pw_reduce")
    dw, db = Guarded((C, 3), torch.float32), Guarded((C,), torch.float32)
    dw.t.zero_(); db.t.zero_()
    _lib.check(lib.pg_pw_wgrad(act.data_ptr(), img.data_ptr(), dw.t.data_ptr(), db.t.data_ptr(), N, H * W, 3, C, 3, 1,
                               ctypes.c_float(1.0), 1, st), "pg_pw_wgrad")
    up = Guarded((N, 2 * H, 2 * W, C), bf)
    _lib.check(lib.pg_upsample2(act.data_ptr(), up.t.data_ptr(), N, H, W, C, 1, st), "pg_upsample2")
    dn = Guarded((N, H, W, C), bf)
    big = torch.randn(N, 2 * H, 2 * W, C, device=DEV).to(bf)
    _lib.check(lib.pg_upsample2_bwd(big.data_ptr(), dn.t.data_ptr(), N, H, W, C, 1, st), "pg_upsample2_bwd")
    r = torch.rand(N, H, W, device=DEV) + 0.5
    da = Guarded((N, H, W, C), bf)
    _lib.check(lib.pg_pn_lrelu_bwd(act.data_ptr(), act.data_ptr(), r.data_ptr(), da.t.data_ptr(), N * H * W, C,
                                   ctypes.c_float(0.2), 1, 0, 0, None, None, 1, st), "pg_pn_lrelu_bwd")
    torch.cuda.synchronize()
    for g_, what in ((out, "pg_pw_expand"), (oimg, "pg_pw_reduce"), (dw, "pg_pw_wgrad(dw)"), (db, "pg_pw_wgrad(db)"),
                     (up, "pg_upsample2"), (dn, "pg_upsample2_bwd"), (da, "pg_pn_lrelu_bwd")):
        g_.check(what)
