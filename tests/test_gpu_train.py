"""GPU: the full iteration (D phase, D Adam, G phase, G Adam, EMA) through
progan_b200.Trainer on the real CUDA kernels."""
import pytest
import torch

import common
import helpers
import progan_b200
from test_trainer import check_against_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(name, precision, impl, use_graph=False, iters=1, segment_graphs=None):
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = impl, impl == "tc"
    K.invalidate_packs()
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, precision, device=DEV)
    Grun, _ = helpers.build_models(inp, precision, device=DEV)
    tr = progan_b200.Trainer(G, D, Grun, use_graph=use_graph, segment_graphs=segment_graphs)
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    for _ in range(iters):
        tr.step(real, z, eps, inp["step"], inp["alpha"])
    torch.cuda.synchronize()
    return inp, tr, G, D, Grun


@pytest.mark.parametrize("name", list(common.CASES))
def test_trainer_fp32_matches_reference_golden(name):
    """fp32 check mode: parameters after both Adam updates, the EMA generator, the three loss
    terms and the G-phase gradients against the golden vectors of the real reference."""
    inp, tr, G, D, Grun = _run(name, "fp32", "simt")
    check_against_golden(name, tr, G, D, Grun, 1e-3)


@pytest.mark.parametrize("name", ["s2_a0.5", "s3_a0.25", "s5_a0.5"])
def test_trainer_bf16_tc_runs_and_graph_equals_eager(name):
    """Product mode: three iterations eager vs three iterations replayed from the captured
    CUDA graph must agree (same kernels, same order), and stay finite."""
    # one iteration: only the fp32 atomics order differs between runs
    _, e1, _, _, _ = _run(name, "bf16", "tc", use_graph=False, iters=1)
    _, g1, _, _, _ = _run(name, "bf16", "tc", use_graph=True, iters=1)
    # (a handful of near-zero gradients may change sign with the atomics order: 2 lr each after Adam)
    assert helpers.rel(g1.bD.p, e1.bD.p) < 1e-4 and helpers.rel(g1.bG.p, e1.bG.p) < 1e-4
    m_e1, m_g1 = e1.read_metrics(), g1.read_metrics()
    for k in m_e1:
        assert abs(m_e1[k] - m_g1[k]) <= 1e-4 * abs(m_e1[k]) + 1e-4, (k, m_e1[k], m_g1[k])
    # three iterations: a single rounding flip (atomics order of the two concurrent gradient
    # chains) is amplified by Adam's normalisation — runs bifurcate at the 1e-4 level whether
    # eager or replayed, so the comparison is loose here
    _, tr_e, Ge, De, _ = _run(name, "bf16", "tc", use_graph=False, iters=3)
    _, tr_g, Gg, Dg, _ = _run(name, "bf16", "tc", use_graph=True, iters=3)
    me, mg = tr_e.read_metrics(), tr_g.read_metrics()
    for k in me:
        assert me[k] == me[k] and abs(me[k]) < 1e6, (k, me[k])          # finite
        assert abs(me[k] - mg[k]) <= 1e-2 * abs(me[k]) + 5e-2, (k, me[k], mg[k])
    rd, rg = helpers.rel(tr_g.bD.p, tr_e.bD.p), helpers.rel(tr_g.bG.p, tr_e.bG.p)
    assert rd < 2e-3, (rd, rg, me, mg)
    assert rg < 2e-3, (rd, rg, me, mg)
    assert float(tr_g.bD.steps.max()) == 3.0
    # the multi-GPU form: three graphs per iteration (all-reduce points between them)
    _, tr_s, _, _, _ = _run(name, "bf16", "tc", use_graph=True, iters=3, segment_graphs=True)
    assert helpers.rel(tr_s.bD.p, tr_e.bD.p) < 2e-3
    assert helpers.rel(tr_s.bG.p, tr_e.bG.p) < 2e-3


def _trainer_grads(name):
    """Gradient buckets of one Trainer iteration, captured right before each Adam step."""
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "tc", True
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, "bf16", device=DEV)
    tr = progan_b200.Trainer(G, D, None)
    snaps, orig = [], tr._adam

    def spy(bucket, plan):
        snaps.append(bucket.g.clone())
        orig(bucket, plan)

    tr._adam = spy
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    tr.step(real, z, eps, inp["step"], inp["alpha"])
    torch.cuda.synchronize()
    return inp, tr, D, snaps


@pytest.mark.parametrize("name", ["s2_a0.5", "s3_a0.25", "s5_a0.5"])
def test_trainer_fast_paths_match_plain_autograd(name):
    """The Trainer's fast paths (one D pass over cat([real, fake]) with per-half minibatch
    statistics, gradients accumulated straight into the flat bucket, deferred weight-gradient
    unpack) give the same D gradients as the reference-ordered step through plain autograd
    (three separate D calls, AccumulateGrad), and are reproducible run to run."""
    inp, tr, D, snaps = _trainer_grads(name)
    _, tr2, _, snaps2 = _trainer_grads(name)
    assert helpers.rel(snaps2[0], snaps[0]) < 1e-5          # only atomics ordering differs
    assert helpers.rel(snaps2[1], snaps[1]) < 1e-5
    G0, D0 = helpers.build_models(inp, "bf16", device=DEV)
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    res, _ = helpers.product_train_step(G0, D0, real, z, eps, inp["step"], inp["alpha"])
    names = {id(p): k for k, p in D.named_parameters()}
    for gname, (a, b) in tr.bD.group_range.items():
        off = a
        for p in tr.bD.group_params[gname]:
            key = names[id(p)]
            if key in res["d_grads"]:
                got = snaps[0][off:off + p.numel()].view(p.shape)
                assert helpers.rel(got, res["d_grads"][key]) < 5e-3, key   # fused bias sums are fp32, autograd sums bf16 dA
            off += p.numel()


def test_n_critic_graph_replay_equals_eager():
    """n_critic = 2: critic-only and full iterations are separate captured graphs (one or
    segmented); four iterations replayed vs eager, and the generator moved exactly twice."""
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "tc", True
    inp = common.make_inputs("s3_a0.25")
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    runs = []
    for use_graph, seg in ((False, None), (True, None), (True, True)):
        K.invalidate_packs()
        G, D = helpers.build_models(inp, "bf16", device=DEV)
        tr = progan_b200.Trainer(G, D, None, use_graph=use_graph, segment_graphs=seg, n_critic=2)
        g0 = tr.bG.p.clone()
        for i in range(4):
            tr.step(real, z, eps, inp["step"], inp["alpha"])
            if i == 0:
                torch.cuda.synchronize()
                assert torch.equal(tr.bG.p, g0)             # critic-only iteration
        torch.cuda.synchronize()
        assert float(tr.bG.steps.max()) == 2.0 and float(tr.bD.steps.max()) == 4.0
        runs.append(tr)
    for other in runs[1:]:
        assert helpers.rel(other.bD.p, runs[0].bD.p) < 2e-3
        assert helpers.rel(other.bG.p, runs[0].bG.p) < 2e-3
        me, mo = runs[0].read_metrics(reset=False), other.read_metrics(reset=False)
        for k in me:
            assert abs(me[k] - mo[k]) <= 1e-2 * abs(me[k]) + 5e-2, (k, me[k], mo[k])


@pytest.mark.parametrize("name", ["s3_a0.25", "s5_a0.5"])
def test_fused_activation_backward_epilogue_in_the_full_iteration(name):
    """The data-gradient convs with the fused PixelNorm/LeakyReLU-backward epilogue (off by
    default) through a whole iteration: the gradient buckets agree with the unfused path (the
    fused path skips one bf16 rounding of dh, hence 2e-2) and the fused entry point was used."""
    K = progan_b200.get_kernels()
    _, off, _, _, _ = _run(name, "bf16", "tc")
    calls = []
    orig = K._call

    def counting(fn, *a):
        calls.append(fn)
        return orig(fn, *a)
    prev = K.fuse_actbwd_max_cout
    K.fuse_actbwd_max_cout, K._call = 128, counting
    try:
        _, on, _, _, _ = _run(name, "bf16", "tc")
    finally:
        K.fuse_actbwd_max_cout, K._call = prev, orig
    assert calls.count("pg_conv_tc_actbwd") >= 2, "the fused kernel did not run"
    assert helpers.rel(on.bD.g, off.bD.g) < 2e-2 and helpers.rel(on.bG.g, off.bG.g) < 2e-2
    mo, mf = off.read_metrics(), on.read_metrics()
    for k in mo:     # gen_loss follows D's Adam step (lr * sign(g) at v = 0), which amplifies the difference
        tol = 1e-2 if k == "gen_loss" else 1e-3
        assert abs(mo[k] - mf[k]) <= tol * abs(mo[k]) + tol, (k, mo[k], mf[k])


@pytest.mark.xfail(reason="added after the round's GPU budget was spent: never yet run on a GPU", strict=False)
def test_g_running_samples_follow_the_ema_under_graph_replay():
    """the captured EMA kernel rewrites g_running's weights through raw pointers,
    so its cached operand copies must be dropped after every replay — a sample drawn after more
    iterations must use the current weights, not the copies cached by the first sample."""
    K = progan_b200.get_kernels()
    inp, tr, G, D, Grun = _run("s3_a0.25", "bf16", "tc", use_graph=True, iters=1)
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    with torch.no_grad():
        first = Grun(z, step=inp["step"], alpha=inp["alpha"]).clone()      # caches operand copies
    for _ in range(3):
        tr.step(real, z, eps, inp["step"], inp["alpha"])
    with torch.no_grad():
        later = Grun(z, step=inp["step"], alpha=inp["alpha"]).clone()
    K.invalidate_packs()                                                   # force fresh copies
    with torch.no_grad():
        fresh = Grun(z, step=inp["step"], alpha=inp["alpha"])
    assert torch.equal(later, fresh), "g_running sampled with stale operand copies"
    assert not torch.equal(first, later), "the EMA did not move g_running"
