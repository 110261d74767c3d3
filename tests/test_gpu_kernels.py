"""GPU unit tests: every C-ABI kernel against the torch specification of its semantics
(tests/emul_kernels.py, running plain fp32 torch ops on the same device).
Tolerances: fp32 kernels 1e-4 relative (l2), bf16 kernels 1e-2 (north_star)."""
import pytest
import torch

import helpers
import progan_b200
from emul_kernels import EmulKernels
from progan_b200.kernels import ConvOp, EPI_LINEAR, EPI_LRELU, EPI_PN_LRELU, CudaKernels

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def KE():
    return CudaKernels(), EmulKernels()


def tol(dtype):
    return 1e-4 if dtype == torch.float32 else 1e-2


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype).contiguous()


DT = [torch.float32, torch.bfloat16]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout, k, pad, swap, flip
    (2, 8, 8, 32, 64, 3, 1, False, False),
    (3, 4, 4, 33, 32, 3, 1, False, False),      # mbstd layer shape (C+1 input channels)
    (4, 4, 4, 32, 32, 4, 0, False, False),      # D last conv 4x4 valid
    (4, 1, 1, 16, 32, 4, 3, True, True),        # G input ConvTranspose 4x4 on 1x1
    (2, 8, 8, 64, 32, 3, 1, True, True),        # data-gradient form
])
@pytest.mark.parametrize("epi", [EPI_LINEAR, EPI_PN_LRELU, EPI_LRELU])
def test_conv_fwd_simt(KE, dtype, cfg, epi):
    K, E = KE
    K.conv_impl = "simt"
    N, H, W, Cin, Cout, k, pad, swap, flip = cfg
    op = ConvOp(k, pad, swap, flip)
    x = rnd(N, H, W, Cin, dtype=dtype)
    w = rnd(*((Cin, Cout, k, k) if swap else (Cout, Cin, k, k)), seed=1)
    b = rnd(Cout, seed=2, scale=0.1)
    scale = (2.0 / (Cin * k * k)) ** 0.5
    y, r = K.conv_fwd(x, w, b, op, scale, epi, 0.2)
    # spec sees the same rounded operands
    wq = w.to(dtype).float() if dtype == torch.bfloat16 else w
    ye, re_ = E.conv_fwd(x, wq, b, op, scale, epi, 0.2)
    assert helpers.rel(y, ye) < tol(dtype)
    if epi == EPI_PN_LRELU:
        assert helpers.rel(r, re_) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", [
    (2, 8, 8, 32, 64, 3, 1, False, False),
    (3, 4, 4, 33, 32, 3, 1, False, False),
    (4, 4, 4, 32, 32, 4, 0, False, False),
    (4, 1, 1, 16, 32, 4, 3, True, True),
    (2, 16, 16, 64, 32, 3, 1, True, True),
])
def test_conv_wgrad_simt(KE, dtype, cfg):
    K, E = KE
    K.conv_impl = "simt"
    N, H, W, Cin, Cout, k, pad, swap, flip = cfg
    op = ConvOp(k, pad, swap, flip)
    Ho = H + 2 * pad - k + 1
    x = rnd(N, H, W, Cin, dtype=dtype)
    dy = rnd(N, Ho, Ho, Cout, dtype=dtype, seed=3)
    wshape = (Cin, Cout, k, k) if swap else (Cout, Cin, k, k)
    dw = K.conv_wgrad(x, dy, wshape, op, 0.37)
    dwe = E.conv_wgrad(x, dy, wshape, op, 0.37)
    assert dw.shape == dwe.shape
    assert helpers.rel(dw, dwe) < 1e-4


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout
    (2, 16, 16, 64, 64), (1, 32, 32, 32, 64), (2, 16, 16, 128, 128), (4, 8, 8, 64, 128),
    (16, 4, 4, 128, 128), (3, 4, 4, 64, 32), (1, 64, 64, 64, 32), (1, 32, 32, 32, 32),
    (2, 32, 32, 128, 64), (5, 8, 8, 32, 128),
    # second-generation kernel: resident weights / streamed weights with two tiles per unit
    (3, 32, 32, 64, 64), (2, 64, 64, 32, 64), (2, 32, 32, 64, 128), (4, 32, 32, 128, 128),
    (1, 64, 32, 64, 32), (3, 32, 64, 32, 32), (1, 32, 32, 128, 32),
    # third-generation kernel: streamed weights, two tiles per stage, cluster multicast (2 and 4)
    (10, 64, 64, 128, 128), (19, 32, 32, 128, 128), (5, 128, 128, 64, 64),
    # wide layers (Correct* defaults): 256-channel N tiles, PixelNorm as a separate kernel above 256
    (2, 16, 16, 512, 512), (3, 8, 8, 256, 512), (1, 32, 32, 512, 256), (5, 4, 4, 512, 512),
])
@pytest.mark.parametrize("epi", [EPI_LINEAR, EPI_PN_LRELU, EPI_LRELU])
@pytest.mark.parametrize("flip", [False, True])
def test_conv_tc_matches_spec(KE, cfg, epi, flip):
    """tcgen05 implicit-GEMM conv vs the torch spec on identical bf16-rounded operands."""
    K, E = KE
    K.conv_impl = "tc"
    N, H, W, Cin, Cout = cfg
    op = ConvOp(3, 1, flip, flip)
    x = rnd(N, H, W, Cin, dtype=torch.bfloat16)
    w = rnd(*((Cin, Cout, 3, 3) if flip else (Cout, Cin, 3, 3)), seed=1)
    b = rnd(Cout, seed=2, scale=0.1)
    scale = (2.0 / (Cin * 9)) ** 0.5
    assert K.tc_eligible(x, w.shape, op)
    y, r = K.conv_fwd(x, w, b, op, scale, epi, 0.2)
    ye, re_ = E.conv_fwd(x, w.to(torch.bfloat16).float(), b, op, scale, epi, 0.2)
    torch.cuda.synchronize()
    err = helpers.rel(y, ye)
    assert err < 5e-3, "rel err %g" % err
    if epi == EPI_PN_LRELU:
        # wide layers: the statistic is taken from the bf16-rounded pre-activation (stand-alone kernel)
        assert helpers.rel(r, re_) < (1e-4 if Cout <= 256 else 1e-3)
    K.conv_impl = "simt"


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout
    (2, 16, 16, 64, 64), (1, 32, 32, 32, 64), (2, 16, 16, 128, 128), (4, 8, 8, 64, 128),
    (16, 4, 4, 128, 128), (3, 4, 4, 64, 32), (1, 64, 64, 64, 32), (1, 32, 32, 32, 32),
    (2, 32, 32, 128, 64), (5, 8, 8, 32, 128), (40, 16, 16, 128, 128),
    # second-generation (column-halo) kernel: every Cin/Cout atom configuration
    (3, 32, 32, 64, 64), (2, 64, 64, 32, 64), (2, 32, 32, 64, 128), (4, 32, 32, 128, 128),
    (1, 64, 32, 64, 32), (3, 32, 64, 32, 32), (1, 32, 32, 128, 32), (2, 32, 32, 32, 128),
    (6, 32, 32, 128, 64),
    # wide layers: dy tiles of 256 channels
    (2, 16, 16, 512, 512), (2, 8, 8, 256, 512), (3, 16, 16, 512, 256), (9, 4, 4, 512, 512),
])
@pytest.mark.parametrize("flip", [False, True])
def test_conv_wgrad_tc_matches_spec(KE, cfg, flip):
    """tcgen05 weight gradient (MN-major operands, TMEM-resident accumulators) vs the spec."""
    K, E = KE
    K.conv_impl, K.wgrad_tc = "tc", True
    N, H, W, Cin, Cout = cfg
    op = ConvOp(3, 1, flip, flip)
    x = rnd(N, H, W, Cin, dtype=torch.bfloat16)
    dy = rnd(N, H, W, Cout, dtype=torch.bfloat16, seed=3)
    wshape = (Cin, Cout, 3, 3) if flip else (Cout, Cin, 3, 3)
    dw = K.conv_wgrad(x, dy, wshape, op, 0.37)
    dwe = E.conv_wgrad(x, dy, wshape, op, 0.37)
    torch.cuda.synchronize()
    err = helpers.rel(dw, dwe)
    K.conv_impl = "simt"
    assert err < 1e-3, "rel err %g" % err


GEMM_CASES = [
    # name, N, H, W, wshape, op, xC (physical x channels), epi
    ("valid_fwd", 64, 4, 4, (128, 128, 4, 4), ConvOp(4, 0), 128, EPI_PN_LRELU),
    ("valid_fwd_small_batch", 6, 4, 4, (64, 32, 4, 4), ConvOp(4, 0), 32, EPI_LRELU),
    ("valid_dgrad(full)", 64, 1, 1, (128, 128, 4, 4), ConvOp(4, 3, True, True), 128, EPI_LINEAR),
    ("convT_fwd(full)", 64, 1, 1, (128, 128, 4, 4), ConvOp(4, 3, True, True), 128, EPI_PN_LRELU),
    ("convT_fwd(full)_z64", 10, 1, 1, (64, 128, 4, 4), ConvOp(4, 3, True, True), 64, EPI_PN_LRELU),
    ("convT_dgrad(valid)", 64, 4, 4, (128, 128, 4, 4), ConvOp(4, 0, False, False), 128, EPI_LINEAR),
    ("mbstd_conv_padded", 64, 4, 4, (128, 129, 3, 3), ConvOp(3, 1, False, False, 160, 0), 160, EPI_PN_LRELU),
    ("mbstd_conv_dgrad_padded", 64, 4, 4, (128, 129, 3, 3), ConvOp(3, 1, True, True, 0, 160), 128, EPI_LINEAR),
    ("mbstd_conv_padded_c32", 5, 4, 4, (32, 33, 3, 3), ConvOp(3, 1, False, False, 64, 0), 64, EPI_PN_LRELU),
    # wide layers (Correct* defaults of 512 channels; z + embedding = 1024 latent channels)
    ("valid_fwd_wide", 8, 4, 4, (512, 512, 4, 4), ConvOp(4, 0), 512, EPI_PN_LRELU),
    ("valid_dgrad(full)_wide", 8, 1, 1, (512, 512, 4, 4), ConvOp(4, 3, True, True), 512, EPI_LINEAR),
    ("convT_fwd(full)_wide", 8, 1, 1, (1024, 512, 4, 4), ConvOp(4, 3, True, True), 1024, EPI_PN_LRELU),
    ("convT_dgrad(valid)_wide", 8, 4, 4, (1024, 512, 4, 4), ConvOp(4, 0, False, False), 512, EPI_LINEAR),
    ("mbstd_conv_padded_wide", 8, 4, 4, (512, 513, 3, 3), ConvOp(3, 1, False, False, 768, 0), 768, EPI_PN_LRELU),
    ("mbstd_conv_dgrad_padded_wide", 8, 4, 4, (512, 513, 3, 3), ConvOp(3, 1, True, True, 0, 768), 512, EPI_LINEAR),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=[c[0] for c in GEMM_CASES])
def test_tc_gemm_and_padded_forms(KE, case):
    """The 4x4 valid conv / 4x4 ConvTranspose-on-1x1 GEMM forms and the channel-padded 3x3 conv
    after minibatch-stddev, forward + weight gradient, vs the spec."""
    K, E = KE
    K.conv_impl, K.wgrad_tc = "tc", True
    name, N, H, W, wshape, op, xC, epi = case
    w = rnd(*wshape, seed=1)
    cin_l, cout_l = op.cin(wshape), op.cout(wshape)
    x = rnd(N, H, W, xC, dtype=torch.bfloat16)
    if xC > cin_l:
        x[..., cin_l:] = 0
    b = rnd(cout_l, seed=2, scale=0.1) if epi != EPI_LINEAR else None
    scale = (2.0 / (cin_l * op.k * op.k)) ** 0.5
    mode = K.tc_mode(x.dtype, H, W, wshape, op)
    assert mode is not None, "expected a tensor-core form"
    y, r = K.conv_fwd(x, w, b, op, scale, epi, 0.2)
    ye, re_ = E.conv_fwd(x, w, b, op, scale, epi, 0.2)
    torch.cuda.synchronize()
    assert y.shape == ye.shape
    assert helpers.rel(y, ye) < 5e-3, (mode, helpers.rel(y, ye))
    if epi == EPI_PN_LRELU:
        assert helpers.rel(r, re_) < (1e-4 if op.cout(wshape) <= 256 else 1e-3)
    dy = rnd(*y.shape, dtype=torch.bfloat16, seed=3)
    if op.ypad:
        dy[..., cout_l:] = 0
    dw = K.conv_wgrad(x, dy, wshape, op, 0.37)
    dwe = E.conv_wgrad(x, dy, wshape, op, 0.37)
    torch.cuda.synchronize()
    K.conv_impl = "simt"
    assert helpers.rel(dw, dwe) < 1e-3, (mode, helpers.rel(dw, dwe))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("P,C", [(70001, 32), (33333, 128), (9999, 512), (100, 2048)])
def test_colsum_long_columns(KE, dtype, P, C):
    """Bias-gradient column sum over many pixels (the unrolled main loop and its remainder)."""
    K, E = KE
    x = rnd(1, 1, P, C, dtype=dtype, seed=31)
    assert helpers.rel(K.colsum(x), E.colsum(x)) < 1e-4
    out = torch.ones(C, device=DEV)
    K.colsum(x, out=out)
    assert helpers.rel(out, E.colsum(x) + 1) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C", [32, 256, 512, 1024])
@pytest.mark.parametrize("use_pn", [True, False])
def test_pn_lrelu_fwd(KE, dtype, C, use_pn):
    """Stand-alone PixelNorm + LeakyReLU forward (layers wider than one conv N tile) vs torch
    (progan_modules.py:58-60, 138), also in place."""
    K, _ = KE
    a = rnd(3, 5, 7, C, seed=21).to(dtype)
    af = a.float()
    rr = torch.rsqrt((af * af).mean(-1) + 1e-8)
    ye = torch.nn.functional.leaky_relu(af * rr.unsqueeze(-1) if use_pn else af, 0.2)
    y, r = K.pn_lrelu_fwd(a, 0.2, use_pn)
    assert helpers.rel(y, ye) < tol(dtype)
    if use_pn:
        assert helpers.rel(r, rr) < 1e-5
    else:
        assert r is None
    y2, _ = K.pn_lrelu_fwd(a, 0.2, use_pn, out=a)
    assert y2.data_ptr() == a.data_ptr() and torch.equal(y2, y)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C", [32, 64, 128, 512])
@pytest.mark.parametrize("use_pn", [True, False])
def test_pn_lrelu_grads(KE, dtype, C, use_pn):
    K, E = KE
    P = (3, 5, 7)
    a = rnd(*P, C, seed=4)
    r = torch.rsqrt((a * a).mean(-1) + 1e-8).contiguous()
    y = torch.nn.functional.leaky_relu(a * r.unsqueeze(-1) if use_pn else a, 0.2).to(dtype)
    dy, t = rnd(*P, C, dtype=dtype, seed=5), rnd(*P, C, dtype=dtype, seed=6)
    da, cs = K.pn_lrelu_bwd(dy, y, r, 0.2, use_pn, False, True)
    dae, cse = E.pn_lrelu_bwd(dy, y, r, 0.2, use_pn, False, True)
    assert helpers.rel(da, dae) < tol(dtype)
    assert helpers.rel(cs, cse) < 1e-3
    # pooled form: dy is the gradient of avgpool2(y) (needs even H, W)
    y2, r2 = y.reshape(3, 5, 7, C)[:, :4, :6].contiguous(), r.reshape(3, 5, 7)[:, :4, :6].contiguous()
    dyp = rnd(3, 2, 3, C, dtype=dtype, seed=7)
    dap, _ = K.pn_lrelu_bwd(dyp, y2, r2, 0.2, use_pn, True, False)
    assert helpers.rel(dap, E.pn_lrelu_bwd(dyp, y2, r2, 0.2, use_pn, True)[0]) < tol(dtype)
    t2 = t.reshape(3, 5, 7, C)[:, :4, :6].contiguous()
    p1, p2 = K.pn_lrelu_bwd_bwd(t2, dyp, y2, r2, 0.2, use_pn, True)
    q1, q2 = E.pn_lrelu_bwd_bwd(t2, dyp, y2, r2, 0.2, use_pn, True)
    assert helpers.rel(p1, q1) < tol(dtype)
    if use_pn:
        assert helpers.rel(p2, q2) < tol(dtype)
    c1, c2 = K.pn_lrelu_bwd_bwd(t, dy, y, r, 0.2, use_pn)
    e1, e2 = E.pn_lrelu_bwd_bwd(t, dy, y, r, 0.2, use_pn)
    assert helpers.rel(c1, e1) < tol(dtype)
    if use_pn:
        assert helpers.rel(c2, e2) < tol(dtype)
    else:
        assert float(c2.float().abs().max()) == 0.0
    assert helpers.rel(K.colsum(dy), E.colsum(dy)) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("Kc,C,HW", [(3, 32, (16, 16)), (3, 128, (8, 8)), (1, 128, (1, 1)), (4, 64, (4, 4)), (3, 512, (8, 8)), (1, 1024, (2, 2)), (3, 264, (4, 4))])
@pytest.mark.parametrize("ck", [True, False])
def test_pointwise_heads(KE, dtype, Kc, C, HW, ck):
    K, E = KE
    N = 5
    w = rnd(C, Kc) if ck else rnd(Kc, C)
    w_sc, w_sk = (Kc, 1) if ck else (1, C)
    img = rnd(N, Kc, *HW, seed=7)
    act = rnd(N, *HW, C, dtype=dtype, seed=8)
    bc, bk = rnd(C, seed=9), rnd(Kc, seed=10)
    assert helpers.rel(K.pw_expand(img, w, bc, C, w_sc, w_sk, 0.7, dtype),
                       E.pw_expand(img, w, bc, C, w_sc, w_sk, 0.7, dtype)) < tol(dtype)
    assert helpers.rel(K.pw_reduce(act, w, bk, Kc, w_sc, w_sk, 0.7),
                       E.pw_reduce(act, w, bk, Kc, w_sc, w_sk, 0.7)) < 1e-4
    assert helpers.rel(K.pw_wgrad(act, img, w.shape, w_sc, w_sk, 0.7),
                       E.pw_wgrad(act, img, w.shape, w_sc, w_sk, 0.7)) < 1e-4
    assert helpers.rel(K.img_chansum(img), E.img_chansum(img)) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape,fmt", [((2, 8, 8, 32), "nhwc"), ((3, 4, 4, 129), "nhwc"),
                                       ((2, 3, 16, 16), "nchw"), ((1, 32, 32, 64), "nhwc")])
def test_resample(KE, dtype, shape, fmt):
    K, E = KE
    if fmt == "nchw" and dtype != torch.float32:
        pytest.skip("images are fp32")
    x = rnd(*shape, dtype=dtype, seed=11)
    for name in ("avgpool2", "avgpool2_bwd", "upsample2", "upsample2_bwd"):
        a, b = getattr(K, name)(x, fmt), getattr(E, name)(x, fmt)
        assert a.shape == b.shape, name
        assert helpers.rel(a, b) < tol(dtype), name
    al = torch.tensor(0.3, device=DEV)
    y = rnd(*shape, dtype=dtype, seed=12)
    assert helpers.rel(K.blend(x, y, al), E.blend(x, y, al)) < tol(dtype)
    assert helpers.rel(K.scale(x, 1.0, -1.0, al), E.scale(x, 1.0, -1.0, al)) < tol(dtype)
    assert helpers.rel(K.scale(x, 0.5, 0.0, None), E.scale(x, 0.5, 0.0, None)) < tol(dtype)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("N,C", [(4, 32), (64, 128), (7, 512)])
def test_mbstd(KE, dtype, N, C):
    K, E = KE
    x = rnd(N, 4, 4, C, dtype=dtype, seed=13)
    Cp = C + 1 if N != 64 else ((C + 1 + 31) // 32) * 32      # also the channel-padded form
    (o, stats), (oe, _) = K.mbstd_fwd(x, Cp), E.mbstd_fwd(x, Cp)
    assert helpers.rel(o, oe) < tol(dtype)
    dout, t = rnd(N, 4, 4, Cp, dtype=dtype, seed=14), rnd(N, 4, 4, C, dtype=dtype, seed=15)
    dout[..., C + 1:] = 0
    assert helpers.rel(K.mbstd_bwd(dout, x, stats), E.mbstd_bwd(dout, x)) < tol(dtype)
    a1, a2 = K.mbstd_bwd_bwd(t, dout, x, stats)
    b1, b2 = E.mbstd_bwd_bwd(t, dout, x)
    assert helpers.rel(a1, b1) < tol(dtype)
    assert helpers.rel(a2, b2) < (1e-3 if dtype == torch.float32 else 2e-2)


def test_gp_and_optim(KE):
    K, E = KE
    N, D = 6, 3 * 16 * 16
    real, fake = rnd(N, 3, 16, 16, seed=16), rnd(N, 3, 16, 16, seed=17)
    eps = torch.rand(N, 1, 1, 1, device=DEV)
    xh = K.interp_xhat(real, fake, eps)
    assert torch.equal(xh, eps * real + (1 - eps) * fake)      # bit-exact eps indexing
    g = rnd(N, 3, 16, 16, seed=18, scale=0.05)
    gp, norms = K.gp_fwd(g, 10.0)
    gpe, ne = E.gp_fwd(g, 10.0)
    assert helpers.rel(gp, gpe) < 1e-5 and helpers.rel(norms, ne) < 1e-5
    up = torch.tensor(0.7, device=DEV)
    assert helpers.rel(K.gp_bwd(g, norms, up, 10.0), E.gp_bwd(g, ne, up, 10.0)) < 1e-5
    # Adam (beta1 = 0) + EMA against torch.optim.Adam
    p = torch.nn.Parameter(rnd(1000, seed=19))
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.0, 0.99))
    p2, v = p.detach().clone(), torch.zeros(1000, device=DEV)
    step = torch.zeros((), device=DEV)
    ema, ema_ref = p.detach().clone(), p.detach().clone()
    for it in range(3):
        grad = rnd(1000, seed=20 + it)
        p.grad = grad.clone()
        opt.step()
        step += 1
        K.adam_step(p2, grad, None, v, 1e-3, 0.0, 0.99, 1e-8, step)
        K.ema(ema, p2, 0.999)
        ema_ref.mul_(0.999).add_(p.detach(), alpha=0.001)
    assert helpers.rel(p2, p) < 1e-6
    assert helpers.rel(ema, ema_ref) < 1e-6
    x = rnd(4, 3, 8, 8, seed=30)
    y = K.tanh_fwd(x)
    assert helpers.rel(y, torch.tanh(x)) < 1e-6
    assert helpers.rel(K.tanh_bwd(x, y), x * (1 - y * y)) < 1e-6


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout  (Cout = channels of the activation whose backward is fused)
    (2, 32, 32, 64, 64), (1, 64, 64, 64, 32), (3, 16, 16, 128, 32), (2, 32, 32, 128, 64),
    (5, 128, 128, 64, 64), (2, 64, 64, 32, 64), (10, 64, 64, 64, 32),
    (3, 16, 16, 128, 128), (10, 64, 64, 128, 128), (4, 32, 32, 64, 128),      # two-pass epilogue
])
@pytest.mark.parametrize("use_pn", [True, False])
def test_conv_dgrad_actbwd_fused_matches_two_kernels(KE, cfg, use_pn):
    """The data-gradient conv with the fused PixelNorm+LeakyReLU backward epilogue (+ bias
    gradient) against the unfused pair pg_conv_tc -> pg_pn_lrelu_bwd on the same operands."""
    K, E = KE
    K.conv_impl = "tc"
    N, H, W, Cin, Cout = cfg
    op = ConvOp(3, 1, True, True)                 # adjoint form, as in a data-gradient
    x = rnd(N, H, W, Cin, dtype=torch.bfloat16)
    w = torch.nn.Parameter(rnd(Cin, Cout, 3, 3, seed=1))
    scale = (2.0 / (Cout * 9)) ** 0.5
    a_prev = rnd(N, H, W, Cout, seed=3)
    r_prev = torch.rsqrt((a_prev ** 2).mean(-1) + 1e-8) if use_pn else None
    pnorm = a_prev * r_prev.unsqueeze(-1) if use_pn else a_prev
    y_prev = torch.where(pnorm > 0, pnorm, 0.2 * pnorm).to(torch.bfloat16)
    cs = torch.zeros(Cout, device=DEV)
    prev_max, K.fuse_actbwd_max_cout = K.fuse_actbwd_max_cout, 128      # the fusion is off by default
    try:
        da = K.conv_dgrad_actbwd(x, w, op, scale, y_prev, r_prev, 0.2, use_pn, cs)
    finally:
        K.fuse_actbwd_max_cout = prev_max
    assert da is not None
    dh, _ = K.conv_fwd(x, w, None, op, scale, EPI_LINEAR)
    cs_ref = torch.zeros(Cout, device=DEV)
    da_ref, _ = K.pn_lrelu_bwd(dh, y_prev, r_prev, 0.2, use_pn, False, False, cs_ref)
    torch.cuda.synchronize()
    # the fused path skips one bf16 rounding (dh stays fp32), so it is the more accurate one
    assert helpers.rel(da, da_ref) < 8e-3
    assert helpers.rel(cs, cs_ref) < 8e-3
    K.conv_impl = "simt"


@pytest.mark.parametrize("cfg", [(2, 32, 32, 64, 64), (3, 16, 16, 128, 128), (1, 64, 64, 32, 64),
                                 (10, 64, 64, 128, 128), (2, 32, 64, 64, 32)])
@pytest.mark.parametrize("epi", [EPI_PN_LRELU, EPI_LRELU])
def test_conv_fused_avgpool_matches_separate_kernel(KE, cfg, epi):
    """The conv epilogue's fused 2x2 average pool (bilinear x0.5, progan_modules.py:299) equals
    pg_avgpool2 of the conv output bit for bit (same bf16 inputs, fp32 sum, one rounding)."""
    K, E = KE
    K.conv_impl = "tc"
    N, H, W, Cin, Cout = cfg
    op = ConvOp(3, 1)
    x = rnd(N, H, W, Cin, dtype=torch.bfloat16)
    w = torch.nn.Parameter(rnd(Cout, Cin, 3, 3, seed=1))
    b = rnd(Cout, seed=2, scale=0.1)
    y, r, yp = K.conv_fwd(x, w, b, op, 0.05, epi, 0.2, pool_out=True)
    assert yp is not None
    y2, r2 = K.conv_fwd(x, w, b, op, 0.05, epi, 0.2)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    assert torch.equal(yp, K.avgpool2(y2))
    K.conv_impl = "simt"


@pytest.mark.parametrize("n_real,n_fake", [(64, 64), (5, 3), (0, 64), (0, 7)])
def test_wgan_loss_seed_matches_autograd(KE, n_real, n_fake):
    """pg_wgan_loss: loss value and dL/d(outputs) of train.py:126-139 / 162-167 vs torch autograd."""
    K, E = KE
    d = rnd(n_real + n_fake, 1, seed=5).requires_grad_(True)
    if n_real:
        dr, df = d[:n_real], d[n_real:]
        real_predict = dr.mean() - 0.001 * (dr ** 2).mean()
        loss, logged = -real_predict + df.mean(), real_predict - df.mean()
    else:
        loss = -d.mean()
        logged = loss
    (g,) = torch.autograd.grad(loss, d)
    metric = torch.full((), 0.25, device=DEV)
    seed = K.wgan_loss(d.detach(), n_real, 0.001, metric)
    torch.cuda.synchronize()
    assert helpers.rel(seed, g) < 1e-6
    assert abs(float(metric) - 0.25 - float(logged)) < 1e-5
