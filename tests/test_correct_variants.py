"""The Correct* rewiring (progan_modules.py:479-598: cifar_train.py / proper_cifar_train.py
models), the mnist_pggan.py models (BASELINE config 0) and the class-conditional Correct* models
(:601-775, BASELINE configs 3 and 5) and the remaining conditional rewirings (:314-476, :778-915,
mnist_pggan.py:140-345: label plane, normalised latent, projection critic) through the product modules, pinned to golden vectors recorded from the REAL
reference classes: forward outputs, gradient penalty, every parameter gradient and the
parameters after torch.optim.Adam(betas=(0, .99)) steps, EMA included.

CPU: kernels emulated in torch (host/autograd logic).  GPU: the CUDA kernels in fp32 check
mode against the same fixtures, and the bf16 tcgen05 path against the fp32 one."""
import os

import pytest
import torch
from torch import optim

import common
import helpers
import progan_b200
from emul_kernels import EmulKernels


def _run_case(name, precision, device):
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, precision, device=device, name=name)
    Grun, _ = helpers.build_models(inp, precision, device=device, name=name)
    g_opt = optim.Adam(G.parameters(), lr=0.001, betas=(0.0, 0.99))
    d_opt = optim.Adam(D.parameters(), lr=0.001, betas=(0.0, 0.99))
    real, z, eps = inp["real"].to(device), inp["z"].to(device), inp["eps"].to(device)
    label = inp["label"].to(device) if inp["label"] is not None else None
    res, fake = helpers.product_train_step(G, D, real, z, eps, inp["step"], inp["alpha"], fused_gp=False,
                                           label=label)
    d_opt.step()
    loss, g_grads = helpers.product_g_phase(G, D, fake, inp["step"], inp["alpha"], label=label)
    g_opt.step()
    with torch.no_grad():
        for (k, pr), (_, pg) in zip(Grun.named_parameters(), G.named_parameters()):
            pr.mul_(0.999).add_(pg, alpha=1 - 0.999)
    return inp, res, loss, g_grads, G, D, Grun


def _check(name, res, loss, g_grads, G, D, Grun, tol):
    gold = torch.load(os.path.join(common.HERE, name + ".pt"), weights_only=True)
    for k in ("real_predict", "fake", "hat_predict", "grad_x_hat"):
        assert helpers.rel(res[k], gold[k]) < tol, k
    assert helpers.rel(res["grad_penalty"], gold["grad_penalty"]) < tol
    assert abs(float(loss) - float(gold["gen_loss"])) <= tol * abs(float(gold["gen_loss"])) + 1e-6
    assert sorted(res["d_grads"]) == sorted(gold["d_grads"])
    for k, g in res["d_grads"].items():
        s = gold["d_grads"][k]
        assert float((common.summarize(g, k) - s).norm()) <= 20 * tol * float(s.norm()) + 1e-7, k
    for k, g in g_grads.items():
        if k in gold["g_grads"]:
            s = gold["g_grads"][k]
            assert float((common.summarize(g, k) - s).norm()) <= 20 * tol * float(s.norm()) + 1e-7, k
    for tag, mod in (("d_params_after", D), ("g_params_after", G), ("g_running_after", Grun)):
        for k, p in mod.named_parameters():
            s = gold[tag][k]
            assert float((common.summarize(p, k) - s).norm()) <= tol * float(s.norm()) + 1e-7, (tag, k)


@pytest.mark.parametrize("name", common.VARIANT_CASES)
def test_correct_variants_match_reference_golden_cpu(name):
    prev = progan_b200.set_kernels(EmulKernels())
    try:
        inp, res, loss, g_grads, G, D, Grun = _run_case(name, "fp32", "cpu")
        _check(name, res, loss, g_grads, G, D, Grun, 2e-4)
    finally:
        progan_b200.set_kernels(prev)


def test_correct_state_dict_layout():
    """Key names/order of the mirrors = what the golden generator asserted against the reference."""
    G = progan_b200.CorrectGenerator(input_code_dim=16, in_channel=32)
    keys = list(G.state_dict().keys())
    assert keys[:4] == ["progression_4.0.conv.bias", "progression_4.0.conv.weight_orig",
                        "progression_4.3.conv.bias", "progression_4.3.conv.weight_orig"]
    assert G.state_dict()["progression_4.0.conv.weight_orig"].shape == (16, 32, 4, 4)     # IOHW
    D = progan_b200.CorrectDiscriminator(feat_dim=32)
    assert D.state_dict()["progression.3.conv.0.conv.weight_orig"].shape == (32, 33, 3, 3)
    assert D.state_dict()["progression.3.conv.3.conv.weight_orig"].shape == (32, 32, 4, 4)
    with pytest.raises(RuntimeError):
        D(torch.zeros(1, 3, 4, 4), step=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", common.VARIANT_CASES)
def test_correct_variants_cuda_check_mode_and_bf16(name):
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "simt", False
    K.invalidate_packs()
    inp, res, loss, g_grads, G, D, Grun = _run_case(name, "fp32", "cuda")
    _check(name, res, loss, g_grads, G, D, Grun, 1e-3)
    # product mode (bf16, tcgen05 where the shapes allow): forward within the north_star gate
    K.conv_impl, K.wgrad_tc = "tc", True
    K.invalidate_packs()
    inp, res16, loss16, _, _, _, _ = _run_case(name, "bf16", "cuda")
    assert helpers.rel(res16["fake"], res["fake"]) < 1e-2
    # D outputs are O(0.1) scalars per sample (batch 2-4): absolute + relative bound
    dp = (res16["real_predict"] - res["real_predict"]).abs().max()
    assert float(dp) <= 2e-2 * float(res["real_predict"].abs().max()) + 2e-2
    assert helpers.rel(res16["grad_x_hat"], res["grad_x_hat"]) < 0.3         # bf16 LeakyReLU mask flips: 3-20 %
    # (||g|| - 1)^2 amplifies the relative error of ||g|| near 1: absolute + relative bound
    assert abs(float(res16["grad_penalty"]) - float(res["grad_penalty"])) <= 0.2 * float(res["grad_penalty"]) + 0.1


@pytest.mark.parametrize("name", ["c2_a0.5", "c4_a0.5", "m2_a0.5", "m3_a0.25", "k2_a0.5_eq", "k3_a0.25", "k5_a0.5",
                                  "b2_a0.5", "b3_a0.25_nopn", "a2_a0.5", "a4_a0.25", "n3_a0.5", "p2_a0.5",
                                  "p3_a1.0_nopn"])
def test_trainer_generic_families_match_golden_cpu(name):
    """progan_b200.Trainer (flat buckets, fused Adam/EMA, one D pass over cat([real, fake]), direct
    gradient accumulation) on the other model families: generic bucket layout + probe for the live
    parameter set, against the goldens of the real reference loop."""
    from test_trainer import check_against_golden
    prev = progan_b200.set_kernels(EmulKernels())
    try:
        inp = common.make_inputs(name)
        G, D = helpers.build_models(inp, "fp32", name=name)
        Grun, _ = helpers.build_models(inp, "fp32", name=name)
        tr = progan_b200.Trainer(G, D, Grun)
        tr.step(inp["real"], inp["z"], inp["eps"], inp["step"], inp["alpha"], label=inp["label"])
        check_against_golden(name, tr, G, D, Grun, 2e-4)
    finally:
        progan_b200.set_kernels(prev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c3_a0.25", "m2_a0.5", "k3_a0.25", "b2_a0.5", "a4_a0.25", "p2_a0.5"])
def test_trainer_generic_families_cuda(name):
    """The same on the CUDA kernels: fp32 check mode against the goldens, and the bf16 product
    path eager vs CUDA-graph replay."""
    from test_trainer import check_against_golden
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = "simt", False
    K.invalidate_packs()
    inp = common.make_inputs(name)
    dev = "cuda"
    lab = inp["label"].to(dev) if inp["label"] is not None else None
    G, D = helpers.build_models(inp, "fp32", device=dev, name=name)
    Grun, _ = helpers.build_models(inp, "fp32", device=dev, name=name)
    tr = progan_b200.Trainer(G, D, Grun)
    tr.step(inp["real"].to(dev), inp["z"].to(dev), inp["eps"].to(dev), inp["step"], inp["alpha"], label=lab)
    check_against_golden(name, tr, G, D, Grun, 1e-3)
    K.conv_impl, K.wgrad_tc = "tc", True
    res = []
    for use_graph in (False, True):
        K.invalidate_packs()
        G, D = helpers.build_models(inp, "bf16", device=dev, name=name)
        tr = progan_b200.Trainer(G, D, None, use_graph=use_graph)
        tr.step(inp["real"].to(dev), inp["z"].to(dev), inp["eps"].to(dev), inp["step"], inp["alpha"], label=lab)
        torch.cuda.synchronize()
        res.append(tr)
    assert helpers.rel(res[1].bD.p, res[0].bD.p) < 1e-5
    assert helpers.rel(res[1].bG.p, res[0].bG.p) < 1e-5


@pytest.mark.gpu
def test_wide_correct_models_tcgen05_vs_cuda_core():
    """The Correct* defaults (512 channels, BASELINE configs 2/3/5): 3x3 convs wider than one
    256-column N tile run as several N tiles with the stand-alone PixelNorm kernel, weight
    gradients over dy tiles of 256 channels.  Same bf16 precision on the CUDA-core kernels is the
    checker (only accumulation order and one extra bf16 rounding of the pre-activation differ)."""
    K = progan_b200.get_kernels()
    dev = "cuda"
    g = torch.Generator().manual_seed(5)
    B, step, alpha = 4, 3, 0.5
    real = (torch.rand(B, 3, 16, 16, generator=g) * 2 - 1).to(dev)
    z = torch.randn(B, 512, generator=g).to(dev)
    eps = torch.rand(B, 1, 1, 1, generator=g).to(dev)
    res = {}
    try:
        for wide in (False, True):
            K.conv_impl, K.wgrad_tc, K.wide_tc = "tc", True, wide
            K.invalidate_packs()
            torch.manual_seed(0)
            G = progan_b200.CorrectGenerator(512, 512, precision="bf16").to(dev)
            D = progan_b200.CorrectDiscriminator(512, precision="bf16").to(dev)
            K.launches = 0
            r, fake = helpers.product_train_step(G, D, real, z, eps, step, alpha)
            loss, g_grads = helpers.product_g_phase(G, D, fake, step, alpha)
            torch.cuda.synchronize()
            res[wide] = (r, g_grads)
    finally:
        K.wide_tc = True
    (a, ga), (b, gb) = res[False], res[True]
    assert helpers.rel(b["fake"], a["fake"]) < 1e-2
    for k in ("real_predict", "hat_predict"):
        assert float((b[k] - a[k]).abs().max()) <= 2e-2 * float(a[k].abs().max()) + 2e-2, k
    assert helpers.rel(b["grad_x_hat"], a["grad_x_hat"]) < 0.3
    assert abs(float(b["grad_penalty"]) - float(a["grad_penalty"])) <= 0.2 * float(a["grad_penalty"]) + 0.1
    for k in a["d_grads"]:
        assert helpers.rel(b["d_grads"][k], a["d_grads"][k]) < 0.3, k
    for k in ga:
        assert helpers.rel(gb[k], ga[k]) < 0.3, k
