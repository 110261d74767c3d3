"""ADA pipe (progan_b200.ada): bit-exact against the reference's AugmentPipe / AdaptiveAugment
where /root/reference exists, against committed golden vectors everywhere; twice differentiable
(the reference's pipe is not on torch >= 1.10); usable in front of the critic in Trainer."""
import os
import sys
import warnings

import pytest
import torch

import progan_b200
from progan_b200 import ada as A

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ada_golden.pt")
CONFIGS = {
    "blit": dict(xflip=1, rotate90=1, xint=1),
    "geom": dict(scale=1, rotate=1, aniso=1, xfrac=1),
    "color": dict(brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1),
    "filter": dict(imgfilter=1),
    "corrupt": dict(noise=1, cutout=1),
    "bgc": dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1,
                lumaflip=1, hue=1, saturation=1),
    "all": dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1,
                lumaflip=1, hue=1, saturation=1, imgfilter=1, noise=1, cutout=1),
}
SHAPES = [(4, 3, 32, 32, 1.0), (3, 1, 32, 32, 0.7), (2, 3, 32, 48, 0.5)]


def _digest(t):
    """Compact fingerprint of an output tensor kept in the fixture: l2 norm, sum, two seeded random
    projections and 512 strided samples."""
    t = t.detach().double().flatten()
    g = torch.Generator().manual_seed(t.numel())
    p1 = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    p2 = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    idx = torch.linspace(0, t.numel() - 1, 512).long()
    return torch.cat([torch.stack([t.norm(), t.sum(), t @ p1, t @ p2]), t[idx]])


def _inputs(shape):
    g = torch.Generator().manual_seed(11)
    return torch.randn(shape[:4], generator=g)


def _reference():
    if not os.path.isdir("/root/reference/ada"):
        return None
    for p in ("/root/reference/ada", "/root/reference"):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from ada import augment as RA
        from ada import adapt_augm as RC
    return RA, RC


def make_golden():
    """Run by `python tests/test_ada.py` in the build container: records the REFERENCE outputs."""
    RA, _ = _reference()
    out = {}
    for name, kw in CONFIGS.items():
        ref = RA.AugmentPipe(**kw)
        for shape in SHAPES[:2]:
            ref.p.fill_(shape[4])
            torch.manual_seed(7)
            out["%s/%s" % (name, "x".join(map(str, shape[:4])))] = _digest(ref(_inputs(shape)))
    torch.save(out, GOLD)
    print("wrote", GOLD, os.path.getsize(GOLD) // 1024, "KiB")


@pytest.mark.parametrize("name", list(CONFIGS))
def test_pipe_matches_reference_golden(name):
    gold = torch.load(GOLD, weights_only=True)
    pipe = progan_b200.AugmentPipe(**CONFIGS[name])
    for shape in SHAPES[:2]:
        pipe.p.fill_(shape[4])
        torch.manual_seed(7)
        d = _digest(pipe(_inputs(shape)))
        ref = gold["%s/%s" % (name, "x".join(map(str, shape[:4])))]
        assert float((d - ref).abs().max()) <= 1e-5 * (1 + float(ref.abs().max())), (name, shape)


@pytest.mark.skipif(_reference() is None, reason="live reference only exists in the build container")
@pytest.mark.parametrize("name", list(CONFIGS))
def test_pipe_is_bit_exact_against_the_live_reference(name):
    RA, _ = _reference()
    ref, mine = RA.AugmentPipe(**CONFIGS[name]), progan_b200.AugmentPipe(**CONFIGS[name])
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert torch.equal(ref.Hz_geom, mine.Hz_geom) and float((ref.Hz_fbank - mine.Hz_fbank).abs().max()) < 1e-8
    for shape in SHAPES:
        ref.p.fill_(shape[4]); mine.p.fill_(shape[4])
        x = _inputs(shape)
        for pct in (None, 0.8, 0.25):
            torch.manual_seed(5); yr = ref(x, debug_percentile=pct)
            torch.manual_seed(5); ym = mine(x, debug_percentile=pct)
            assert torch.equal(yr, ym), (name, shape, pct, float((yr - ym).abs().max()))


@pytest.mark.skipif(_reference() is None, reason="live reference only exists in the build container")
def test_adaptive_augment_follows_the_reference_controller():
    _, RC = _reference()
    g = torch.Generator().manual_seed(3)
    ref, mine = RC.AdaptiveAugment(0.1, 0.6, 2000, 8), progan_b200.AdaptiveAugment(0.1, 0.6, 2000, 8)
    assert ref.initialize() == mine.initialize()
    for i in range(40):
        logits = torch.randn(8, 1, generator=g) + (0.8 if i < 20 else -0.5)
        assert abs(ref.update(logits) - mine.update(logits)) < 1e-12
    ref.set_batch_size(16); mine.set_batch_size(16)
    assert ref.ada_aug_step == mine.ada_aug_step


def test_sampled_parameters_can_be_replayed():
    pipe = progan_b200.AugmentPipe(**CONFIGS["all"])
    x = _inputs((2, 3, 32, 32))
    torch.manual_seed(1)
    P = pipe.sample(2, 3, 32, 32, x.device)
    y1, y2 = pipe.apply(x, P), pipe.apply(x, P)
    assert torch.equal(y1, y2)
    torch.manual_seed(1)
    assert torch.equal(pipe(x), y1)


def test_pipe_is_affine_in_the_images_and_twice_differentiable():
    """The property the gradient penalty needs: d/dx and d2/dx2 through every stage.  In float64,
    with one fixed parameter draw, (a) apply() is affine, (b) gradcheck / gradgradcheck pass, and
    (c) a WGAN-GP style double backward through the pipe runs (the reference's pipe raises
    'derivative for aten::grid_sampler_2d_backward is not implemented' here)."""
    pipe = progan_b200.AugmentPipe(**CONFIGS["all"]).double()
    torch.manual_seed(2)
    P = pipe.sample(2, 3, 32, 32, torch.device("cpu"))
    g = torch.Generator().manual_seed(4)
    x1 = torch.randn(2, 3, 32, 32, generator=g, dtype=torch.float64)
    x2 = torch.randn(2, 3, 32, 32, generator=g, dtype=torch.float64)
    f = lambda x: pipe.apply(x, P)                                     # noqa: E731
    lhs, rhs = f(0.3 * x1 + 0.7 * x2), 0.3 * f(x1) + 0.7 * f(x2)
    assert float((lhs - rhs).abs().max()) < 1e-10
    small = AugSmall()
    assert torch.autograd.gradcheck(small, (small.x0.clone().requires_grad_(True),), eps=1e-6, atol=1e-6)
    assert torch.autograd.gradgradcheck(small, (small.x0.clone().requires_grad_(True),), eps=1e-6, atol=1e-6)
    x = x1.clone().requires_grad_(True)
    w = torch.randn(3, 32, 32, generator=g, dtype=torch.float64) * 0.01     # keeps tanh out of saturation
    score = torch.tanh((f(x) * w).flatten(1).sum(1))                    # a stand-in critic
    (gx,) = torch.autograd.grad(score.sum(), x, create_graph=True)
    gp = ((gx.flatten(1).norm(dim=1) - 1) ** 2).mean()
    gp.backward()
    assert x.grad is not None and bool(torch.isfinite(x.grad).all()) and float(x.grad.abs().sum()) > 0


class AugSmall:
    """A small fixed-parameter instance for gradcheck (8x8 would be below the sym2 filter bank's
    reflect padding, so the band filter is left out here and covered by the affine check above)."""

    def __init__(self):
        kw = dict(CONFIGS["bgc"], noise=1, cutout=1)
        self.pipe = progan_b200.AugmentPipe(**kw).double()
        torch.manual_seed(9)
        self.P = self.pipe.sample(1, 3, 8, 8, torch.device("cpu"))
        self.x0 = torch.randn(1, 3, 8, 8, generator=torch.Generator().manual_seed(6), dtype=torch.float64)

    def __call__(self, x):
        return torch.tanh(self.pipe.apply(x, self.P))


def test_trainer_runs_with_the_pipe_in_front_of_the_critic():
    """One eager iteration with ADA in front of every critic input (kernel emulation on the CPU):
    the gradient penalty differentiates through the pipe, losses stay finite, and the critic's
    gradients differ from the un-augmented run."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import common
    import helpers
    from emul_kernels import EmulKernels
    prev = progan_b200.set_kernels(EmulKernels())
    try:
        inp = common.make_inputs("s3_a0.25")
        grads = []
        for aug in (None, progan_b200.AugmentPipe(**CONFIGS["bgc"])):
            G, D = helpers.build_models(inp, "fp32")
            tr = progan_b200.Trainer(G, D, None, augment=aug)
            torch.manual_seed(0)
            tr.step(inp["real"], inp["z"], inp["eps"], inp["step"], inp["alpha"])
            m = tr.read_metrics()
            assert all(v == v and abs(v) < 1e6 for v in m.values()), m
            grads.append(tr.bD.g.clone())
        assert float((grads[0] - grads[1]).abs().max()) > 0
        with pytest.raises(RuntimeError):
            progan_b200.Trainer(G, D, None, augment=aug, use_graph=True)
    finally:
        progan_b200.set_kernels(prev)


@pytest.mark.gpu
@pytest.mark.xfail(reason="written after the round's GPU budget was spent: never yet run on a GPU", strict=False)
def test_pipe_on_the_gpu_between_x_hat_and_the_critic():
    dev = "cuda"
    pipe = progan_b200.AugmentPipe(**CONFIGS["all"]).to(dev)
    cpu = progan_b200.AugmentPipe(**CONFIGS["all"])
    x = _inputs((4, 3, 32, 32))
    torch.manual_seed(1)
    P = cpu.sample(4, 3, 32, 32, torch.device("cpu"))
    Pd = A.AugmentParams(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(P).items()})
    assert float((pipe.apply(x.to(dev), Pd).cpu() - cpu.apply(x, P)).abs().max()) < 1e-4
    D = progan_b200.Discriminator(32).to(dev)
    xh = x.to(dev).requires_grad_(True)
    hat = D(pipe(xh).contiguous(), step=3, alpha=0.5)
    (g,) = torch.autograd.grad(hat.sum(), xh, create_graph=True)
    progan_b200.gradient_penalty(g, 10.0).backward()
    assert all(bool(torch.isfinite(p.grad).all()) for p in D.parameters() if p.grad is not None)


if __name__ == "__main__":
    make_golden()
