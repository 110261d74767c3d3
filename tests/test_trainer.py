"""Full iteration (train.py:97-169 with Adam(0.0, 0.99), EMA 0.999) through
progan_b200.Trainer on the CPU test double, pinned to the golden vectors recorded from the
REAL reference including both optimiser updates and the EMA."""
import os

import pytest
import torch

import common
import helpers
import progan_b200
from emul_kernels import EmulKernels


@pytest.fixture(autouse=True)
def emul_backend():
    prev = progan_b200.set_kernels(EmulKernels())
    yield
    progan_b200.set_kernels(prev)


def run_trainer_case(name, precision, device="cpu", use_graph=False, iters=1):
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, precision, device=device)
    Grun, _ = helpers.build_models(inp, precision, device=device)
    tr = progan_b200.Trainer(G, D, Grun, use_graph=use_graph)
    real, z, eps = inp["real"].to(device), inp["z"].to(device), inp["eps"].to(device)
    for _ in range(iters):
        tr.step(real, z, eps, inp["step"], inp["alpha"])
    return inp, tr, G, D, Grun


def check_against_golden(name, tr, G, D, Grun, tol):
    gold = torch.load(os.path.join(common.HERE, name + ".pt"), weights_only=True)
    m = tr.read_metrics()
    assert abs(m["grad_penalty"] - float(gold["grad_penalty"])) <= tol * abs(float(gold["grad_penalty"]))
    assert abs(m["gen_loss"] - float(gold["gen_loss"])) <= tol * abs(float(gold["gen_loss"])) + 1e-6
    assert abs(m["disc_loss"] - float(gold["disc_loss"])) <= tol * abs(float(gold["disc_loss"])) + 1e-6
    for tag, mod in (("d_params_after", D), ("g_params_after", G), ("g_running_after", Grun)):
        for k, p in mod.named_parameters():
            got = common.summarize(p, k)
            s = gold[tag][k]
            assert float((got - s).norm()) <= tol * float(s.norm()) + 1e-7, (tag, k)
    # G gradients of the G phase (taken with the UPDATED D, as the reference does)
    for k, p in G.named_parameters():
        if k in gold["g_grads"]:
            got = common.summarize(p.grad, k)
            s = gold["g_grads"][k]
            assert float((got - s).norm()) <= 20 * tol * float(s.norm()) + 1e-7, ("g_grads", k)
        else:
            assert k in gold["g_grad_none"] and float(p.grad.abs().max()) == 0.0, k


@pytest.mark.parametrize("name", [n for n in common.CASES if not n.startswith("s5")])
def test_trainer_iteration_matches_golden(name):
    inp, tr, G, D, Grun = run_trainer_case(name, "fp32")
    check_against_golden(name, tr, G, D, Grun, 2e-4)


def test_inactive_parameters_untouched_and_steps_per_group():
    inp, tr, G, D, Grun = run_trainer_case("s1_a1.0", "fp32", iters=2)
    # step 1, no fading: only linear, progression.6/.5 and from_rgb.5 are live in D
    live = {"linear", "progression.6", "progression.5", "from_rgb.5"}
    for name, gi in tr.bD.group_index.items():
        assert float(tr.bD.steps[gi]) == (2.0 if name in live else 0.0), name
    before = common.make_inputs("s1_a1.0")["D"]
    for k, p in D.named_parameters():
        grp = ".".join(k.split(".")[:2]) if not k.startswith("linear") else "linear"
        if grp not in live:
            assert torch.equal(p.detach(), before[k]), k
    assert tr.iterations == 2


@pytest.mark.parametrize("name,n_critic", [("s2_a0.5", 2), ("s3_a0.25", 3), ("k2_a0.5_eq", 2)])
def test_n_critic_matches_reference_ordered_loop(name, n_critic):
    """n_critic > 1 (train.py:158,221): the generator phase, its Adam step and the EMA run only on
    iterations with (i + 1) % n_critic == 0; checked against the loop written with plain autograd
    calls and torch.optim.Adam, four iterations."""
    from torch import optim
    inp = common.make_inputs(name)
    lab = inp["label"]
    G, D = helpers.build_models(inp, "fp32", name=name)
    Grun, _ = helpers.build_models(inp, "fp32", name=name)
    tr = progan_b200.Trainer(G, D, Grun, n_critic=n_critic)
    G2, D2 = helpers.build_models(inp, "fp32", name=name)
    Grun2, _ = helpers.build_models(inp, "fp32", name=name)
    g_opt = optim.Adam(G2.parameters(), lr=0.001, betas=(0.0, 0.99))
    d_opt = optim.Adam(D2.parameters(), lr=0.001, betas=(0.0, 0.99))
    gen = torch.Generator().manual_seed(77)
    gen_loss_sum, n_g = 0.0, 0
    for i in range(4):
        real = torch.rand(inp["real"].shape, generator=gen) * 2 - 1
        z = torch.randn(inp["z"].shape, generator=gen)
        eps = torch.rand(inp["eps"].shape, generator=gen)
        tr.step(real, z, eps, inp["step"], inp["alpha"], label=lab)
        res, fake = helpers.product_train_step(G2, D2, real, z, eps, inp["step"], inp["alpha"], label=lab)
        d_opt.step()
        if (i + 1) % n_critic == 0:
            loss, _ = helpers.product_g_phase(G2, D2, fake, inp["step"], inp["alpha"], label=lab)
            g_opt.step()
            with torch.no_grad():
                for (_, pr), (_, pg) in zip(Grun2.named_parameters(), G2.named_parameters()):
                    pr.mul_(0.999).add_(pg, alpha=1 - 0.999)
            gen_loss_sum += float(loss)
            n_g += 1
    assert tr.iterations == 4 and n_g == 4 // n_critic
    for a, b in ((D, D2), (G, G2), (Grun, Grun2)):
        for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
            assert helpers.rel(p, q) < 1e-5, k
    m = tr.read_metrics()
    assert abs(m["gen_loss"] - gen_loss_sum) <= 1e-4 * abs(gen_loss_sum) + 1e-6
    with pytest.raises(ValueError):
        progan_b200.Trainer(G, D, Grun, n_critic=0)


def test_reduce_ranges_one_aligned_span_minus_what_the_backward_sweep_already_reduced():
    from progan_b200.train import _reduce_ranges
    assert _reduce_ranges([(0, 100), (400, 420)]) == [(0, 420)]              # one collective, gaps are zeros
    assert _reduce_ranges([(6, 101)], total=1000) == [(4, 104)]             # 16-byte boundaries
    assert _reduce_ranges([(6, 999)], total=1000) == [(4, 1000)]
    assert _reduce_ranges([(0, 100), (400, 420)], done=(0, 64)) == [(64, 420)]
    assert _reduce_ranges([(0, 100), (400, 420)], done=(0, 100)) == [(400, 420)]
    assert _reduce_ranges([(0, 100)], done=(0, 100)) == []


def test_tensor_alpha_is_evaluated_like_a_number_unless_marked():
    """a plain tensor alpha outside [0, 1) must not take the blend path."""
    from progan_b200.progan_modules import _fading
    assert _fading(torch.tensor(-1.0)) == (False, -1.0)
    assert _fading(torch.tensor(1.0)) == (False, 1.0)
    assert _fading(torch.tensor(0.25))[0] is True
    t = torch.tensor(0.0)
    t._pg_fading = True
    f, a = _fading(t)
    assert f is True and a is t
    assert _fading(0.5) == (True, 0.5) and _fading(-1) == (False, -1)
