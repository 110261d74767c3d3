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


def _run(name, precision, impl, use_graph=False, iters=1):
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = impl, impl == "tc"
    K.invalidate_packs()
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, precision, device=DEV)
    Grun, _ = helpers.build_models(inp, precision, device=DEV)
    tr = progan_b200.Trainer(G, D, Grun, use_graph=use_graph)
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
    _, tr_e, Ge, De, _ = _run(name, "bf16", "tc", use_graph=False, iters=3)
    _, tr_g, Gg, Dg, _ = _run(name, "bf16", "tc", use_graph=True, iters=3)
    me, mg = tr_e.read_metrics(), tr_g.read_metrics()
    for k in me:
        assert me[k] == me[k] and abs(me[k]) < 1e6, (k, me[k])          # finite
        assert abs(me[k] - mg[k]) <= 5e-2 * abs(me[k]) + 1e-3, (k, me[k], mg[k])
    # atomics in the weight-gradient reductions make runs differ in the last bits only
    assert helpers.rel(tr_g.bD.p, tr_e.bD.p) < 1e-2
    assert helpers.rel(tr_g.bG.p, tr_e.bG.p) < 1e-2
    assert float(tr_g.bD.steps.max()) == 3.0
