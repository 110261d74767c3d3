"""Shared between make_golden.py (runs the REAL reference, only where /root/reference
exists) and the tests (which never need the reference): seeded inputs and the compact
summaries stored in the fixtures."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name: (channel, z_dim, step, alpha, batch, tanh, pixel_norm)
CASES = {
    "s1_a1.0": (32, 32, 1, 1.0, 4, False, True),
    "s1_a0.5": (32, 32, 1, 0.5, 4, False, True),
    "s2_a0.5": (32, 32, 2, 0.5, 4, False, True),
    "s2_a1.0_tanh": (32, 32, 2, 1.0, 4, True, True),
    "s3_a0.25": (32, 32, 3, 0.25, 4, False, True),
    "s3_a-1_nopn": (32, 32, 3, -1, 4, False, False),
    "s5_a0.5": (32, 16, 5, 0.5, 2, False, True),
}


# the Correct* rewiring of the same blocks (progan_modules.py:479-598; step 1 = 4 px ... 4 = 32 px)
CORRECT_CASES = {
    "c1_a1.0": (32, 32, 1, 1.0, 4, False, True),
    "c2_a0.5": (32, 32, 2, 0.5, 4, False, True),
    "c2_a0.5_tanh": (32, 32, 2, 0.5, 4, True, True),      # quirk: tanh path skips the blend at step 2
    "c3_a0.25": (32, 16, 3, 0.25, 4, False, True),
    "c4_a1.0_nopn": (32, 32, 4, 1.0, 2, False, False),
    "c4_a0.5": (32, 32, 4, 0.5, 2, False, True),
}


# mnist_pggan.py (BASELINE config 0): 1-channel images, single-conv blocks, step 1 = 8 px ... 3 = 32 px
MNIST_CASES = {
    "m1_a1.0": (32, 32, 1, 1.0, 4, True, True),
    "m2_a0.5": (32, 32, 2, 0.5, 4, True, True),
    "m3_a0.25": (32, 16, 3, 0.25, 4, False, True),
    "m3_a1.0_nopn": (32, 32, 3, 1.0, 2, True, False),
}


# class-conditional Correct* models (progan_modules.py:601-775; BASELINE configs 3 and 5):
# name: (channel, z_dim, step, alpha, batch, tanh, pixel_norm, num_classes, do_equal_embed)
COND_CASES = {
    "k1_a1.0": (32, 16, 1, 1.0, 4, False, True, 10, False),
    "k2_a0.5_eq": (32, 16, 2, 0.5, 4, False, True, 10, True),
    "k3_a0.25": (32, 16, 3, 0.25, 4, True, True, 14, False),
    "k4_a1.0_eq": (32, 32, 4, 1.0, 2, False, True, 10, True),
    "k5_a0.5": (32, 16, 5, 0.5, 2, False, True, 10, False),
}


# the remaining conditional rewirings; same tuple layout as COND_CASES (do_equal_embed unused)
# progan_modules.ConditionalGenerator / ConditionalDiscriminatorWgangp (:314-476): step 0 = 4 px
CONDBASE_CASES = {
    "b1_a1.0": (32, 16, 1, 1.0, 4, False, True, 10, False),
    "b2_a0.5": (32, 16, 2, 0.5, 4, True, True, 10, False),
    "b3_a0.25_nopn": (32, 32, 3, 0.25, 2, False, False, 7, False),
}
# progan_modules.ConditionalCorrectGeneratorAda / ConditionalCorrectDiscriminatorAda (:778-915)
ADA_CASES = {
    "a1_a1.0": (32, 16, 1, 1.0, 4, False, True, 10, False),
    "a2_a0.5": (32, 16, 2, 0.5, 4, False, True, 10, False),
    "a2_a0.5_tanh": (32, 16, 2, 0.5, 4, True, True, 10, False),
    "a4_a0.25": (32, 32, 4, 0.25, 2, False, True, 14, False),
}
# mnist_pggan.ConditionalGenerator with ConditionalDiscriminatorWgangp / ConditionalDiscriminatorAda
MCOND_CASES = {
    "n1_a1.0": (32, 16, 1, 1.0, 4, True, True, 10, False),
    "n3_a0.5": (32, 16, 3, 0.5, 2, True, True, 10, False),
}
MADA_CASES = {
    "p2_a0.5": (32, 16, 2, 0.5, 4, True, True, 10, False),
    "p3_a1.0_nopn": (32, 32, 3, 1.0, 2, False, False, 10, False),
}

# family -> (cases, generator class, critic class, lives in mnist_pggan, first resolution, image channels)
FAMILIES = {
    "base": (CASES, "Generator", "Discriminator", False, 8, 3),
    "correct": (CORRECT_CASES, "CorrectGenerator", "CorrectDiscriminator", False, 4, 3),
    "mnist": (MNIST_CASES, "Generator", "Discriminator", True, 8, 1),
    "cond": (COND_CASES, "ConditionalCorrectGenerator", "ConditionalCorrectDiscriminatorWgangp", False, 4, 3),
    "condbase": (CONDBASE_CASES, "ConditionalGenerator", "ConditionalDiscriminatorWgangp", False, 8, 3),
    "ada": (ADA_CASES, "ConditionalCorrectGeneratorAda", "ConditionalCorrectDiscriminatorAda", False, 4, 3),
    "mcond": (MCOND_CASES, "ConditionalGenerator", "ConditionalDiscriminatorWgangp", True, 8, 1),
    "mada": (MADA_CASES, "ConditionalGenerator", "ConditionalDiscriminatorAda", True, 8, 1),
}
VARIANT_CASES = [n for f, v in FAMILIES.items() if f != "base" for n in v[0]]
ALL_CASES = list(CASES) + VARIANT_CASES


def family(name):
    for fam, v in FAMILIES.items():
        if name in v[0]:
            return fam
    raise KeyError(name)


def classes(mod, name):
    """(Generator class, Discriminator class) of `mod` — the reference's progan_modules (the
    mnist families live in its sibling module mnist_pggan) or the mirror package."""
    _, g, d, in_mnist, _, _ = FAMILIES[family(name)]
    if in_mnist:
        if hasattr(mod, "mnist_pggan"):
            mod = mod.mnist_pggan
        else:
            import mnist_pggan as mod
    return getattr(mod, g), getattr(mod, d)


def build(mod, name, inp, **extra):
    """(G, D) instances of the case's model family from `mod` (reference module or mirror)."""
    GC, DC = classes(mod, name)
    fam = family(name)
    gk = dict(input_code_dim=inp["z_dim"], in_channel=inp["channel"], pixel_norm=inp["pixel_norm"],
              tanh=inp["tanh"])
    dk = dict(feat_dim=inp["channel"])
    if inp["num_classes"]:
        gk["num_of_classes"] = dk["num_of_classes"] = inp["num_classes"]
    if fam == "cond":
        gk.update(max_step=6, do_equal_embed=inp["equal_embed"])
        dk.update(do_equal_embed=inp["equal_embed"])
    return GC(**gk, **extra), DC(**dk, **extra)


def model_shapes(name, inp):
    """state-dict key -> shape, taken from the host mirror (identical to the reference's;
    make_golden.py asserts that)."""
    import progan_b200
    G, D = build(progan_b200, name, inp)
    return ({k: tuple(v.shape) for k, v in G.state_dict().items()},
            {k: tuple(v.shape) for k, v in D.state_dict().items()})


def make_state(shapes, seed):
    """weight_orig ~ N(0,1) (progan_modules.py:68), bias ~ 0.1*N(0,1) so bias paths are live."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        t = torch.randn(shapes[k], generator=g)
        out[k] = t * 0.1 if k.endswith("bias") else t
    return out


def make_inputs(name):
    fam = family(name)
    cases, _, _, _, res0, img_ch = FAMILIES[fam]
    spec = cases[name]
    ch, zd, step, alpha, B, tanh, pn = spec[:7]
    ncls, eq = (spec[7], spec[8]) if len(spec) > 7 else (0, False)
    inp = dict(step=step, alpha=alpha, tanh=tanh, pixel_norm=pn, channel=ch, z_dim=zd, num_classes=ncls,
               equal_embed=eq)
    gs, ds = model_shapes(name, inp)
    G_state, D_state = make_state(gs, 100), make_state(ds, 200)
    g = torch.Generator().manual_seed(1234)
    R = (res0 // 2) * 2 ** step
    real = torch.rand(B, img_ch, R, R, generator=g) * 2 - 1
    z = torch.randn(B, zd, generator=g)
    eps = torch.rand(B, 1, 1, 1, generator=g)
    label = torch.randint(0, ncls, (B,), generator=g) if ncls else None     # drawn last: older fixtures unchanged
    inp.update(G=G_state, D=D_state, real=real, z=z, eps=eps, label=label)
    return inp


def summarize(t, key):
    """[l2 norm, sum, <t,p1>, <t,p2>] with projections seeded by the key."""
    t = t.detach().double().flatten().cpu()
    g = torch.Generator().manual_seed(abs(hash_str(key)) % (2 ** 31))
    p1 = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    p2 = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    return torch.stack([t.norm(), t.sum(), t @ p1, t @ p2])


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 1000000007
    return h


def summarize_dict(d):
    return {k: summarize(v, k) for k, v in d.items()}
