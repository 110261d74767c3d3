"""GPU: size-independent properties of the hot-path kernels at BASELINE config-4's FULL sizes
(batch 64, 128 px / 64 px, 32..128 channels), where the torch oracle is too slow/large to be the
checker.  Each test states its tolerance; exact properties are checked bit for bit.

  * homogeneity with a power-of-two factor (exact in bf16 and fp32)     -> bit-exact
  * <conv(x;W), y> == <x, conv^T(y;W)>  (data-gradient is the adjoint)   -> 5e-3 (bf16 outputs)
  * <wgrad(x,dy), V> == <conv(x;V), dy>  (weight-gradient is the adjoint) -> 5e-3
  * sum_c p_c * da_c == 0 for the PixelNorm backward (projection)        -> 2e-2 of the norms
  * x_hat = eps*x + (1-eps)*G(z)                                          -> bit-exact vs torch
  * one full Trainer iteration: finite losses, every live parameter moved by at most lr*(1+1e-3)
    (Adam with beta1 = 0 normalises the first step to lr * sign(g))
"""
import pytest
import torch

import progan_b200
from progan_b200.kernels import ConvOp, EPI_LINEAR, CudaKernels

pytestmark = pytest.mark.gpu
DEV = "cuda"
B = 64


@pytest.fixture(scope="module")
def K():
    k = CudaKernels()
    k.conv_impl, k.wgrad_tc = "tc", True
    return k


def _rand(shape, seed, dtype=torch.bfloat16, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


def _dot(a, b):
    return float((a.double() * b.double()).sum())


FULL = [(128, 64, 64), (128, 32, 64), (64, 128, 128), (64, 64, 128), (128, 64, 32)]   # res, Cin, Cout


@pytest.mark.parametrize("res,cin,cout", FULL)
def test_conv_is_homogeneous_bit_exact(K, res, cin, cout):
    x = _rand((B, res, res, cin), 1)
    w = torch.nn.Parameter(_rand((cout, cin, 3, 3), 2, torch.float32))
    op = ConvOp(3, 1)
    y1, _ = K.conv_fwd(x, w, None, op, 0.0625, EPI_LINEAR)
    y2, _ = K.conv_fwd(x * 0.5, w, None, op, 0.0625, EPI_LINEAR)
    assert torch.equal(y2, y1 * 0.5)


@pytest.mark.parametrize("res,cin,cout", FULL)
def test_data_gradient_is_the_adjoint(K, res, cin, cout):
    x = _rand((B, res, res, cin), 3)
    y = _rand((B, res, res, cout), 4)
    w = torch.nn.Parameter(_rand((cout, cin, 3, 3), 5, torch.float32))
    op = ConvOp(3, 1)
    cx, _ = K.conv_fwd(x, w, None, op, 0.05, EPI_LINEAR)
    cty, _ = K.conv_fwd(y, w, None, op.adjoint(), 0.05, EPI_LINEAR)
    lhs, rhs = _dot(cx, y), _dot(x, cty)
    scale = (float(cx.double().norm()) * float(y.double().norm()))
    assert abs(lhs - rhs) <= 5e-3 * scale, (lhs, rhs, scale)     # bf16 rounding of the two outputs


@pytest.mark.parametrize("res,cin,cout", FULL)
def test_weight_gradient_is_the_adjoint(K, res, cin, cout):
    x = _rand((B, res, res, cin), 6)
    dy = _rand((B, res, res, cout), 7)
    v = torch.nn.Parameter(_rand((cout, cin, 3, 3), 8, torch.float32).to(torch.bfloat16).float())
    op = ConvOp(3, 1)
    dw = K.conv_wgrad(x, dy, (cout, cin, 3, 3), op, 1.0)
    cv, _ = K.conv_fwd(x, v, None, op, 1.0, EPI_LINEAR)
    lhs, rhs = _dot(dw, v), _dot(cv, dy)
    scale = float(cv.double().norm()) * float(dy.double().norm())
    assert abs(lhs - rhs) <= 5e-3 * scale, (lhs, rhs, scale)


@pytest.mark.parametrize("res,c", [(128, 64), (64, 128), (128, 32)])
@pytest.mark.parametrize("pool", [False, True])
def test_pixelnorm_backward_is_a_projection(K, res, c, pool):
    a = _rand((B, res, res, c), 9, torch.float32)
    r = torch.rsqrt((a * a).mean(-1) + 1e-8)
    p = a * r.unsqueeze(-1)
    y = torch.where(p > 0, p, 0.2 * p).to(torch.bfloat16)
    dy = _rand((B, res // 2, res // 2, c) if pool else (B, res, res, c), 10)
    da, _ = K.pn_lrelu_bwd(dy, y, r, 0.2, True, pool)
    yf = y.float()
    pf = torch.where(yf > 0, yf, yf / 0.2)
    resid = (pf * da.float()).sum(-1)
    ref = pf.norm(dim=-1) * da.float().norm(dim=-1) + 1e-20
    assert float((resid.abs() / ref).max()) < 2e-2              # bf16 rounding of y and da
    assert float((resid.abs() / ref).mean()) < 3e-3


def test_xhat_bit_exact_full_size(K):
    g = torch.Generator().manual_seed(11)
    real = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(DEV)
    fake = torch.randn(B, 3, 128, 128, generator=g).to(DEV)
    eps = torch.rand(B, 1, 1, 1, generator=g).to(DEV)
    got = K.interp_xhat(real, fake, eps.reshape(-1))
    assert torch.equal(got, eps * real + (1 - eps) * fake)      # train.py:143


def test_full_iteration_first_adam_step_is_lr_sign(K):
    prev = progan_b200.set_kernels(K)
    try:
        torch.manual_seed(0)
        G = progan_b200.Generator(128, 128, tanh=False).to(DEV)
        D = progan_b200.Discriminator(128).to(DEV)
        tr = progan_b200.Trainer(G, D, None, lr=1e-3)
        p0d, p0g = tr.bD.p.clone(), tr.bG.p.clone()
        g = torch.Generator().manual_seed(1234)
        real = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(DEV)
        z = torch.randn(B, 128, generator=g).to(DEV)
        eps = torch.rand(B, 1, 1, 1, generator=g).to(DEV)
        tr.step(real, z, eps, 5, 0.5)
        m = tr.read_metrics()
        assert all(v == v and abs(v) < 1e6 for v in m.values()), m
        for b_, p0 in ((tr.bD, p0d), (tr.bG, p0g)):
            d = (b_.p - p0).abs()
            assert float(d.max()) <= 1e-3 * (1 + 1e-3)
            live = b_.steps > 0
            assert int(live.sum()) >= 7
        # every live D parameter with a non-zero gradient moved by ~lr
        moved = (tr.bD.p - p0d).abs() > 0.9e-3
        assert float(moved.float().mean()) > 0.5
    finally:
        progan_b200.set_kernels(prev)
