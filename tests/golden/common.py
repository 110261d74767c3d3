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


def family(name):
    if name in CORRECT_CASES:
        return "correct"
    if name in MNIST_CASES:
        return "mnist"
    return "cond" if name in COND_CASES else "base"


def build(mod, name, inp, **extra):
    """(G, D) instances of the case's model family from `mod` (reference module or mirror)."""
    GC, DC = classes(mod, name)
    if family(name) == "cond":
        G = GC(input_code_dim=inp["z_dim"], num_of_classes=inp["num_classes"], in_channel=inp["channel"],
               pixel_norm=inp["pixel_norm"], tanh=inp["tanh"], max_step=6,
               do_equal_embed=inp["equal_embed"], **extra)
        D = DC(feat_dim=inp["channel"], num_of_classes=inp["num_classes"],
               do_equal_embed=inp["equal_embed"], **extra)
        return G, D
    G = GC(input_code_dim=inp["z_dim"], in_channel=inp["channel"], pixel_norm=inp["pixel_norm"],
           tanh=inp["tanh"], **extra)
    return G, DC(feat_dim=inp["channel"], **extra)


def classes(mod, name):
    """(Generator class, Discriminator class) of `mod` — the reference's progan_modules (the
    mnist family lives in its sibling module mnist_pggan) or the mirror package."""
    fam = family(name)
    if fam == "correct":
        return mod.CorrectGenerator, mod.CorrectDiscriminator
    if fam == "cond":
        return mod.ConditionalCorrectGenerator, mod.ConditionalCorrectDiscriminatorWgangp
    if fam == "mnist":
        if hasattr(mod, "mnist_pggan"):
            return mod.mnist_pggan.Generator, mod.mnist_pggan.Discriminator
        import mnist_pggan
        return mnist_pggan.Generator, mnist_pggan.Discriminator
    return mod.Generator, mod.Discriminator


def model_shapes(channel, z_dim, pixel_norm, fam="base", ncls=10, eq=False):
    """state-dict key -> shape, taken from the host mirror (identical to the reference's;
    make_golden.py asserts that)."""
    import progan_b200
    if fam in ("correct", "mnist", "cond"):
        if fam == "cond":
            G = progan_b200.ConditionalCorrectGenerator(z_dim, ncls, channel, pixel_norm=pixel_norm,
                                                        max_step=6, do_equal_embed=eq)
            D = progan_b200.ConditionalCorrectDiscriminatorWgangp(channel, ncls, do_equal_embed=eq)
        elif fam == "correct":
            G = progan_b200.CorrectGenerator(input_code_dim=z_dim, in_channel=channel, pixel_norm=pixel_norm)
            D = progan_b200.CorrectDiscriminator(feat_dim=channel)
        else:
            G = progan_b200.mnist_pggan.Generator(input_code_dim=z_dim, in_channel=channel, pixel_norm=pixel_norm)
            D = progan_b200.mnist_pggan.Discriminator(feat_dim=channel)
        return ({k: tuple(v.shape) for k, v in G.state_dict().items()},
                {k: tuple(v.shape) for k, v in D.state_dict().items()})
    G = progan_b200.Generator(input_code_dim=z_dim, in_channel=channel, pixel_norm=pixel_norm)
    D = progan_b200.Discriminator(feat_dim=channel)
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
    spec = CASES.get(name) or CORRECT_CASES.get(name) or MNIST_CASES.get(name) or COND_CASES[name]
    ch, zd, step, alpha, B, tanh, pn = spec[:7]
    ncls, eq = (spec[7], spec[8]) if len(spec) > 7 else (0, False)
    gs, ds = model_shapes(ch, zd, pn, family(name), ncls, eq)
    G_state, D_state = make_state(gs, 100), make_state(ds, 200)
    g = torch.Generator().manual_seed(1234)
    R = 2 * 2 ** step if family(name) in ("correct", "cond") else 4 * 2 ** step
    real = torch.rand(B, 1 if family(name) == "mnist" else 3, R, R, generator=g) * 2 - 1
    z = torch.randn(B, zd, generator=g)
    eps = torch.rand(B, 1, 1, 1, generator=g)
    label = torch.randint(0, ncls, (B,), generator=g) if ncls else None     # drawn last: older fixtures unchanged
    return dict(G=G_state, D=D_state, real=real, z=z, eps=eps, step=step, alpha=alpha,
                tanh=tanh, pixel_norm=pn, channel=ch, z_dim=zd, label=label, num_classes=ncls,
                equal_embed=eq)


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
