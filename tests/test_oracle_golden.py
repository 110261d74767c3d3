"""Pin the oracle: oracle/progan_oracle.py must reproduce the golden vectors that
tests/golden/make_golden.py recorded from the REAL reference (progan_modules.py +
the loop body of train.py:122-169).  Where /root/reference exists (build container) the
oracle is additionally compared tensor-by-tensor against the live reference."""
import os
import sys

import pytest
import torch

import common
import helpers
from oracle import progan_oracle as O


def _close(a, b, tol, what):
    assert helpers.rel(a, b) < tol, what


@pytest.mark.parametrize("name", list(common.CASES) + list(common.CORRECT_CASES))
def test_oracle_matches_golden(name):
    """train.py's models (s*) and the Correct* rewiring of config 2 (c*, recorded from the real
    CorrectGenerator / CorrectDiscriminator with proper_cifar_train.py's loop body)."""
    gold = torch.load(os.path.join(common.HERE, name + ".pt"), weights_only=True)
    inp = common.make_inputs(name)
    step, alpha = inp["step"], inp["alpha"]
    PG, PD, PGrun = O.params_of(inp["G"]), O.params_of(inp["D"]), O.params_of(inp["G"], False)
    optG, optD = O.AdamState(PG), O.AdamState(PD)
    res = O.train_iteration(PG, PD, PGrun, optG, optD, inp["real"], inp["z"], inp["eps"], step,
                            alpha, inp["tanh"], inp["pixel_norm"], family=common.family(name))
    for k in ("real_predict", "fake", "hat_predict", "grad_x_hat", "grad_penalty", "disc_loss",
              "gen_loss"):
        _close(res[k], gold[k], 2e-5, k)
    assert sorted(set(inp["D"]) - set(res["d_grads"])) == gold["d_grad_none"]
    assert sorted(set(inp["G"]) - set(res["g_grads"])) == gold["g_grad_none"]
    for k, s in gold["d_grads"].items():
        _close(common.summarize(res["d_grads"][k], k), s, 2e-4, "d_grads." + k)
    for k, s in gold["g_grads"].items():
        _close(common.summarize(res["g_grads"][k], k), s, 2e-4, "g_grads." + k)
    for k, s in gold["d_params_after"].items():
        _close(common.summarize(PD[k], k), s, 1e-5, "d_after." + k)
    for k, s in gold["g_params_after"].items():
        _close(common.summarize(PG[k], k), s, 1e-5, "g_after." + k)
    for k, s in gold["g_running_after"].items():
        _close(common.summarize(PGrun[k], k), s, 1e-5, "ema." + k)


@pytest.mark.skipif(not os.path.exists("/root/reference/progan_modules.py"),
                    reason="live reference only exists in the build container")
def test_oracle_matches_live_reference_taps():
    sys.path.insert(0, "/root/reference")
    import progan_modules as R
    inp = common.make_inputs("s3_a0.25")
    G = R.Generator(input_code_dim=inp["z_dim"], in_channel=inp["channel"], tanh=False)
    D = R.Discriminator(feat_dim=inp["channel"])
    G.load_state_dict(inp["G"]); D.load_state_dict(inp["D"])
    taps = {}
    hooks = []
    for n, m in list(D.named_modules()) + list(G.named_modules()):
        if isinstance(m, torch.nn.LeakyReLU):
            hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: taps.__setitem__(n, o.detach())))
    with torch.no_grad():
        d_ref = D(inp["real"], step=3, alpha=0.25)
        d_taps = dict(taps); taps.clear()
        g_ref = G(inp["z"], step=3, alpha=0.25)
        g_taps = dict(taps)
        ot, og = {}, {}
        d_or = O.d_forward(O.params_of(inp["D"], False), inp["real"], 3, 0.25, taps=ot)
        g_or = O.g_forward(O.params_of(inp["G"], False), inp["z"], 3, 0.25, tanh=False, taps=og)
    assert helpers.rel(d_or, d_ref) < 1e-6 and helpers.rel(g_or, g_ref) < 1e-6
    # LeakyReLU outputs of the reference sit at conv.2 / conv.5 of each block
    n_cmp = 0
    for k, v in ot.items():
        if k.startswith("progression"):
            blk, idx = k.rsplit(".conv.", 1)
            ref_name = "%s.conv.%d" % (blk, int(idx) + 2)
            assert helpers.rel(v, d_taps[ref_name]) < 1e-6, k
            n_cmp += 1
    for k, v in og.items():
        if k.startswith("progression"):
            blk, idx = k.rsplit(".conv.", 1)
            assert helpers.rel(v, g_taps["%s.conv.%d" % (blk, int(idx) + 2)]) < 1e-6, k
            n_cmp += 1
    assert n_cmp >= 10
    for h in hooks:
        h.remove()
