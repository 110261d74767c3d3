"""Host logic on CPU: the product's modules + autograd Function families, driven through
the torch emulation of the kernels (tests/emul_kernels.py) in fp64, must reproduce the
oracle (and hence the reference, see test_oracle_golden.py) to ~1e-9 — forward,
first-order grads, and the WGAN-GP double backward with the hand-derived PixelNorm and
minibatch-stddev second-order terms."""
import pytest
import torch

import common
import helpers
import progan_b200
from emul_kernels import EmulKernels
from oracle import progan_oracle as O


@pytest.fixture(autouse=True)
def emul_backend():
    prev = progan_b200.set_kernels(EmulKernels())
    yield
    progan_b200.set_kernels(prev)


def _to64(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


@pytest.mark.parametrize("name", [n for n in common.CASES if not n.startswith("s5")])
def test_train_step_matches_oracle_fp64(name):
    inp = common.make_inputs(name)
    step, alpha = inp["step"], inp["alpha"]
    G, D = helpers.build_models(inp, "fp64", dtype=torch.float64)
    real, z, eps = inp["real"].double(), inp["z"].double(), inp["eps"].double()
    res, fake = helpers.product_train_step(G, D, real, z, eps, step, alpha)
    gen_loss, g_grads = helpers.product_g_phase(G, D, fake, step, alpha)

    PG, PD = O.params_of(_to64(inp["G"])), O.params_of(_to64(inp["D"]))
    ref, rfake = O.train_step(PG, PD, real, z, eps, step, alpha, inp["tanh"], inp["pixel_norm"])
    rloss, rg = O.g_phase(PG, PD, rfake, step, alpha)

    tol = 1e-9
    for k in ("real_predict", "fake", "hat_predict", "grad_x_hat", "grad_penalty", "disc_loss"):
        assert helpers.rel(res[k], ref[k]) < tol, k
    assert set(res["d_grads"]) == set(ref["d_grads"])
    for k, v in ref["d_grads"].items():
        assert helpers.rel(res["d_grads"][k], v) < 1e-8, k
    assert helpers.rel(gen_loss, rloss) < tol
    assert set(g_grads) == set(rg)
    for k, v in rg.items():
        assert helpers.rel(g_grads[k], v) < 1e-8, k


@pytest.mark.parametrize("B", [1, 3])
@pytest.mark.parametrize("name", ["s2_a0.5", "s3_a-1_nopn"])
def test_edge_batch_sizes_match_oracle_fp64(name, B):
    """Batch 1 (minibatch-stddev of a single sample: variance 0, sigma = sqrt(1e-8)) and an odd
    batch, first- and second-order, against the oracle."""
    inp = common.make_inputs(name)
    step, alpha = inp["step"], inp["alpha"]
    g = torch.Generator().manual_seed(5 + B)
    real = (torch.rand(B, *inp["real"].shape[1:], generator=g) * 2 - 1).double()
    z = torch.randn(B, inp["z_dim"], generator=g).double()
    eps = torch.rand(B, 1, 1, 1, generator=g).double()
    G, D = helpers.build_models(inp, "fp64", dtype=torch.float64)
    res, fake = helpers.product_train_step(G, D, real, z, eps, step, alpha)
    gen_loss, g_grads = helpers.product_g_phase(G, D, fake, step, alpha)
    PG, PD = O.params_of(_to64(inp["G"])), O.params_of(_to64(inp["D"]))
    ref, rfake = O.train_step(PG, PD, real, z, eps, step, alpha, inp["tanh"], inp["pixel_norm"])
    rloss, rg = O.g_phase(PG, PD, rfake, step, alpha)
    for k in ("real_predict", "fake", "hat_predict", "grad_x_hat", "grad_penalty", "disc_loss"):
        assert helpers.rel(res[k], ref[k]) < 1e-9, k
    for k, v in ref["d_grads"].items():
        assert helpers.rel(res["d_grads"][k], v) < 1e-7, k
    assert helpers.rel(gen_loss, rloss) < 1e-9
    for k, v in rg.items():
        assert helpers.rel(g_grads[k], v) < 1e-7, k


def test_unfused_gp_chain_also_works():
    """The reference script's own torch expression for the penalty (train.py:148-150) must
    keep working on the product modules (drop-in)."""
    inp = common.make_inputs("s2_a0.5")
    G, D = helpers.build_models(inp, "fp64", dtype=torch.float64)
    real, z, eps = inp["real"].double(), inp["z"].double(), inp["eps"].double()
    a, _ = helpers.product_train_step(G, D, real, z, eps, 2, 0.5, fused_gp=True)
    b, _ = helpers.product_train_step(G, D, real, z, eps, 2, 0.5, fused_gp=False)
    assert helpers.rel(a["grad_penalty"], b["grad_penalty"]) < 1e-12
    for k in a["d_grads"]:
        assert helpers.rel(a["d_grads"][k], b["d_grads"][k]) < 1e-10, k


def test_weight_grads_skipped_when_not_requested():
    """autograd.grad(inputs=[x_hat]) must not run weight-gradient kernels (the reference's
    ATen conv backward skips them through its output mask)."""
    inp = common.make_inputs("s1_a1.0")
    G, D = helpers.build_models(inp, "fp64", dtype=torch.float64)
    x = inp["real"].double().requires_grad_(True)
    K = progan_b200.get_kernels()
    calls = []
    orig = K.conv_wgrad
    K.conv_wgrad = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    out = D(x, step=1, alpha=1.0)
    torch.autograd.grad(out.sum(), x, create_graph=True)
    assert not calls
    out.sum().backward()
    assert calls


def test_generator_step0_returns_none_and_state_dict_keys():
    G = progan_b200.Generator(input_code_dim=8, in_channel=32)
    assert G(torch.zeros(2, 8), step=0) is None
    keys = list(G.state_dict().keys())
    assert keys[0] == "input_layer.0.conv.bias" and keys[1] == "input_layer.0.conv.weight_orig"
    D = progan_b200.Discriminator(feat_dim=32)
    assert "linear.linear.weight_orig" in D.state_dict()
    assert D.state_dict()["progression.6.conv.0.conv.weight_orig"].shape == (32, 33, 3, 3)
    assert D.state_dict()["progression.6.conv.3.conv.weight_orig"].shape == (32, 32, 4, 4)


def test_progressive_schedule_matches_reference_loop():
    """ProgressiveSchedule.next() against a literal transcription of train.py:97-111."""
    import progan_b200
    total_iter, init_step = 90, 1
    sched = progan_b200.ProgressiveSchedule(total_iter, init_step=init_step, max_step=3)
    step, iteration = init_step, 0
    for i in range(3 * total_iter):
        alpha = min(1, (2 / (total_iter // 3)) * iteration)          # train.py:100
        reload_data = False
        if iteration > total_iter // 3:                               # train.py:102-111
            alpha = 0
            iteration = 0
            step += 1
            if step > 3:
                alpha = 1
                step = 3
            reload_data = True
        iteration += 1
        assert sched.next() == (step, alpha, reload_data), i
        assert sched.resolution == 4 * 2 ** step


@pytest.mark.parametrize("init_step,max_step,per_mini", [(1, 4, 5), (2, 3, 4), (1, 2, 3), (1, 1, 2)])
def test_mini_step_schedule_matches_reference_loop(init_step, max_step, per_mini):
    """MiniStepSchedule.next() against a literal transcription of proper_cifar_train.py:157-189
    (the `np.inf` sentinel included)."""
    import progan_b200
    batch = 16
    sched = progan_b200.MiniStepSchedule(per_mini * batch + 3, batch, init_step=init_step, max_step=max_step)
    iterations_per_mini_step = (per_mini * batch + 3) // batch              # :77
    step, step_iteration = init_step, 0                                      # :157
    inf = float("inf")
    for i in range(2 * per_mini * (max_step + 2)):
        alpha = min(1, step_iteration / iterations_per_mini_step)            # :165
        reload_data = False
        if step_iteration == iterations_per_mini_step and step == 1:         # :167-171
            alpha = 0
            step_iteration = 0
            step += 1
            reload_data = True
        elif step_iteration == 2 * iterations_per_mini_step:                 # :172-180
            alpha = 0
            step_iteration = 0
            step += 1
            if step > max_step:
                alpha = 1
                step_iteration = inf
                step = max_step
            reload_data = True
        if step_iteration != inf:                                            # :188-189
            step_iteration += 1
        assert sched.next() == (step, alpha, reload_data), i
        assert sched.resolution == 2 * 2 ** step


def test_partial_flush_of_weight_gradient_workspaces_by_address_range(monkeypatch):
    """CudaKernels.flush_wgrads(ptr_range=...) — what the data-parallel Trainer uses to fold and
    all-reduce the top of the critic during the backward sweep: only the workspaces whose gradient
    lives in the range are unpacked and removed from the pending set; two workspaces of the SAME
    gradient go to different launches (the kernel's read-modify-write needs no atomics)."""
    from progan_b200 import _lib
    from progan_b200.kernels import CudaKernels
    K = CudaKernels()                      # the C-ABI library loads without a GPU; nothing is launched
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    calls = []
    K._call = lambda name, *a: calls.append((name, a[1]))         # (entry point, number of table rows)
    K._upload = lambda rows, cls, dev: torch.zeros(len(rows))
    K._stream = lambda: 0

    def pend(grad_ptr, variant):
        ws = torch.zeros(4)
        key = (grad_ptr, variant, False, False, 32, 32, 32, 32, 3, 1.0)
        K._pending[key] = (ws, _lib.UnpackEntry(0, grad_ptr, 32, 32, 32, 32, 9, 0, 0, 0, 1.0, 0.0))
        return key

    early = [pend(1000, "conv3"), pend(1000, "adjoint"), pend(2000, "conv3")]
    late = [pend(5000, "conv3"), pend(6000, "conv3")]
    K.flush_wgrads(ptr_range=(0, 4000), join=False)
    assert [n for n, _ in calls] == ["pg_wgrad_unpack_multi"] * 2 and sorted(r for _, r in calls) == [1, 2]
    assert list(K._pending) == late
    calls.clear()
    K.flush_wgrads(ptr_range=(0, 4000), join=False)               # nothing left in the range
    assert calls == []
    K.flush_wgrads(join=False)
    assert [r for _, r in calls] == [2] and not K._pending
    assert early[0] != early[1]


def test_out_of_range_tensor_alpha_takes_the_no_blend_path():
    """alpha passed as a plain 0-dim tensor of -1 or >= 1 must behave like the number
    (the reference evaluates `0 <= alpha < 1` on it), not like 'fade active'."""
    import common
    import helpers
    import progan_b200
    from emul_kernels import EmulKernels
    prev = progan_b200.set_kernels(EmulKernels())
    try:
        inp = common.make_inputs("s2_a0.5")
        G, D = helpers.build_models(inp, "fp32")
        with torch.no_grad():
            for a in (-1.0, 1.0, 0.5):
                assert torch.equal(D(inp["real"], step=2, alpha=torch.tensor(a)), D(inp["real"], step=2, alpha=a))
                assert torch.equal(G(inp["z"], step=2, alpha=torch.tensor(a)), G(inp["z"], step=2, alpha=a))
            assert not torch.equal(D(inp["real"], step=2, alpha=torch.tensor(-1.0)), D(inp["real"], step=2, alpha=0.5))
    finally:
        progan_b200.set_kernels(prev)
