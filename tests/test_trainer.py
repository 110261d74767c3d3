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
