"""ORACLE — test infrastructure only.  Never imported by the product package.

A plain PyTorch (fp32/fp64, any device) functional restatement of the reference's
progressive-GAN training step, used as the checker for the CUDA path:

  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import this file;
  * pinned against the real reference: tests/golden/make_golden.py imports
    /root/reference/progan_modules.py, runs it on seeded inputs and commits the outputs
    under tests/golden/; tests/test_oracle_golden.py checks this restatement against
    those vectors (the reference itself has no tests or golden vectors — SURVEY.md §4).

Each function cites the reference lines it follows.  Parameters are passed as a dict
{state-dict key: tensor} with the reference's own key names.
"""
from math import sqrt

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ primitives
def eq_scale(w):
    """EqualLR.compute_weight, progan_modules.py:22-27: sqrt(2/fan_in) with
    fan_in = size(1) * numel(w[0][0])  (so ConvTranspose2d uses Cout*k*k)."""
    fan_in = w.size(1) * w[0][0].numel()
    return sqrt(2 / fan_in)


def pixel_norm(x):
    """PixelNorm.forward, progan_modules.py:58-60."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def up2(x):
    """upscale(), progan_modules.py:167-168."""
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def down2(x):
    """progan_modules.py:299,303 (bilinear x0.5 == 2x2 average)."""
    return F.interpolate(x, scale_factor=0.5, mode='bilinear', align_corners=False)


def eq_conv2d(P, prefix, x, padding=0):
    """EqualConv2d.forward, progan_modules.py:63-73 (+ pre-hook :43-45)."""
    w = P[prefix + '.conv.weight_orig']
    return F.conv2d(x, w * eq_scale(w), P[prefix + '.conv.bias'], padding=padding)


def eq_conv_transpose2d(P, prefix, x):
    """EqualConvTranspose2d.forward, progan_modules.py:76-92 (4x4, stride 1, pad 0)."""
    w = P[prefix + '.conv.weight_orig']
    return F.conv_transpose2d(x, w * eq_scale(w), P[prefix + '.conv.bias'], stride=1, padding=0)


def conv_block(P, prefix, x, pixelnorm=True, pad2=1, taps=None, taps_in=None):
    """ConvBlock.forward, progan_modules.py:120-148.  Second conv is 4x4/pad0 for D's last block.
    taps: the activation after each conv+[PN]+LReLU; taps_in: the input each conv saw (for
    teacher-forced per-layer comparisons)."""
    i2 = 3 if pixelnorm else 2
    if taps_in is not None:
        taps_in[prefix + '.conv.0'] = x
    a = eq_conv2d(P, prefix + '.conv.0', x, 1)
    if pixelnorm:
        a = pixel_norm(a)
    h = F.leaky_relu(a, 0.2)
    if taps is not None:
        taps[prefix + '.conv.0'] = h
    if taps_in is not None:
        taps_in[prefix + '.conv.%d' % i2] = h
    a = eq_conv2d(P, prefix + '.conv.%d' % i2, h, pad2)
    if pixelnorm:
        a = pixel_norm(a)
    h = F.leaky_relu(a, 0.2)
    if taps is not None:
        taps[prefix + '.conv.%d' % i2] = h
    return h


# ------------------------------------------------------------------ generator
_G_BLOCKS = ['progression_8', 'progression_16', 'progression_32', 'progression_64',
             'progression_128', 'progression_256']
_G_HEADS = ['to_rgb_8', 'to_rgb_16', 'to_rgb_32', 'to_rgb_64', 'to_rgb_128', 'to_rgb_256']


def g_forward(P, z, step=0, alpha=-1, tanh=True, pixelnorm=True, max_step=6, taps=None, taps_in=None):
    """Generator.forward, progan_modules.py:219-254 (+ progress :204-207, output :209-217)."""
    step = min(step, max_step)
    if step < 1:
        return None
    zdim = P['input_layer.0.conv.weight_orig'].shape[0]
    a = eq_conv_transpose2d(P, 'input_layer.0', z.view(-1, zdim, 1, 1))
    feat = F.leaky_relu(pixel_norm(a), 0.2)          # input_layer always normalises (:181-184)
    if taps is not None:
        taps['input_layer'] = feat
    if taps_in is not None:
        taps_in['input_layer'] = z.view(-1, zdim, 1, 1)
    feat = conv_block(P, 'progression_4', feat, pixelnorm, taps=taps, taps_in=taps_in)
    prev = None
    for s in range(1, step + 1):
        prev = feat
        feat = conv_block(P, _G_BLOCKS[s - 1], up2(feat), pixelnorm, taps=taps, taps_in=taps_in)
    out = eq_conv2d(P, _G_HEADS[step - 1], feat)
    if step >= 2 and 0 <= alpha < 1:                 # no blend at step 1 (:231-234)
        skip = up2(eq_conv2d(P, _G_HEADS[step - 2], prev))
        out = (1 - alpha) * skip + alpha * out
    if tanh:
        out = torch.tanh(out)
    return out


# -------------------------------------------------------------- discriminator
def d_forward(P, x, step=0, alpha=-1, n_layer=7, taps=None, taps_in=None):
    """Discriminator.forward, progan_modules.py:282-311."""
    out = None
    for i in range(step, -1, -1):
        index = n_layer - i - 1
        if i == step:
            out = eq_conv2d(P, 'from_rgb.%d' % index, x)
            if taps is not None:
                taps['from_rgb.%d' % index] = out
        if i == 0:
            out_std = torch.sqrt(out.var(0, unbiased=False) + 1e-8)          # :290
            mean_std = out_std.mean().expand(out.size(0), 1, 4, 4)           # :291-292
            out = torch.cat([out, mean_std], 1)                              # :293
        out = conv_block(P, 'progression.%d' % index, out, True,
                         pad2=0 if index == n_layer - 1 else 1, taps=taps, taps_in=taps_in)
        if i > 0:
            out = down2(out)                                                 # :299
            if i == step and 0 <= alpha < 1:
                skip = eq_conv2d(P, 'from_rgb.%d' % (index + 1), down2(x))   # :303-304
                out = (1 - alpha) * skip + alpha * out                       # :305
    out = out.squeeze(2).squeeze(2)
    w = P['linear.linear.weight_orig']
    return F.linear(out, w * eq_scale(w), P['linear.linear.bias'])            # :307-309



# ------------------------------------------------- Correct* rewiring (config 2)
_C_BLOCKS = ['progression_8', 'progression_16', 'progression_32']
_C_HEADS = ['to_rgb_4', 'to_rgb_8', 'to_rgb_16', 'to_rgb_32']


def correct_g_forward(P, z, step=0, alpha=-1, tanh=False, pixelnorm=True, max_step=4, taps=None,
                      taps_in=None):
    """CorrectGenerator.forward, progan_modules.py:523-545 (output(): :512-521): step 1 = 4 px;
    the 4x4 stem (ConvTranspose + conv, :487-494) always applies PixelNorm; with tanh the
    step-2 path returns without the blend (:534-537)."""
    step = min(step, max_step)
    zdim = P['progression_4.0.conv.weight_orig'].shape[0]
    if taps_in is not None:
        taps_in['progression_4.0'] = z.view(-1, zdim, 1, 1)
    h = F.leaky_relu(pixel_norm(eq_conv_transpose2d(P, 'progression_4.0', z.view(-1, zdim, 1, 1))), 0.2)
    if taps is not None:
        taps['progression_4.0'] = h
    if taps_in is not None:
        taps_in['progression_4.3'] = h
    feat = F.leaky_relu(pixel_norm(eq_conv2d(P, 'progression_4.3', h, 1)), 0.2)
    if taps is not None:
        taps['progression_4.3'] = feat
    if step < 1:
        return None
    prev = None
    for s in range(2, step + 1):
        prev = feat
        feat = conv_block(P, _C_BLOCKS[s - 2], up2(feat), pixelnorm, taps=taps, taps_in=taps_in)
    out = eq_conv2d(P, _C_HEADS[step - 1], feat)
    if step >= 2 and 0 <= alpha < 1 and not (step == 2 and tanh):
        out = (1 - alpha) * up2(eq_conv2d(P, _C_HEADS[step - 2], prev)) + alpha * out
    return torch.tanh(out) if tanh else out


def correct_d_forward(P, x, step=0, alpha=-1, n_layer=4, taps=None, taps_in=None):
    """CorrectDiscriminator.forward, progan_modules.py:576-598: step 1 = 4 px, loop
    range(step, 0, -1), minibatch-stddev before the last block, fade-in at every step > 1."""
    out = None
    for i in range(step, 0, -1):
        index = n_layer - i
        if i == step:
            out = eq_conv2d(P, 'from_rgb.%d' % index, x)
            if taps is not None:
                taps['from_rgb.%d' % index] = out
        if i == 1:
            out_std = torch.sqrt(out.var(0, unbiased=False) + 1e-8)
            out = torch.cat([out, out_std.mean().expand(out.size(0), 1, 4, 4)], 1)
        out = conv_block(P, 'progression.%d' % index, out, True,
                         pad2=0 if index == n_layer - 1 else 1, taps=taps, taps_in=taps_in)
        if i > 1:
            out = down2(out)
            if i == step and 0 <= alpha < 1:
                skip = eq_conv2d(P, 'from_rgb.%d' % (index + 1), down2(x))
                out = (1 - alpha) * skip + alpha * out
    out = out.squeeze(2).squeeze(2)
    w = P['linear.linear.weight_orig']
    return F.linear(out, w * eq_scale(w), P['linear.linear.bias'])


FAMILY_FORWARDS = {
    'base': (g_forward, d_forward),
    'correct': (correct_g_forward, correct_d_forward),
}

# ------------------------------------------------------------------ train step
def params_of(module_or_dict, requires_grad=True, device=None, dtype=None):
    """Detached leaf copies of a module's parameters, keyed like its state_dict."""
    items = module_or_dict.items() if isinstance(module_or_dict, dict) \
        else module_or_dict.named_parameters()
    out = {}
    for k, v in items:
        t = v.detach().clone()
        if device is not None or dtype is not None:
            t = t.to(device=device or t.device, dtype=dtype or t.dtype)
        out[k] = t.requires_grad_(requires_grad)
    return out


def _acc(store, P, keys=None):
    for k, p in P.items():
        if p.grad is not None:
            store[k] = p.grad.clone() if k not in store else store[k] + p.grad
            p.grad = None


def train_step(PG, PD, real, z, eps, step, alpha, tanh=False, pixelnorm=True,
               gp_lambda=10.0, want_taps=False, family='base'):
    """One iteration of the hot loop, train.py:122-167, without the optimiser updates.

    Returns a dict with the three loss terms, x_hat, grad_x_hat, the D parameter gradients
    accumulated over the three D-phase backward calls (train.py:130,139,151), and the G
    parameter gradients of the G phase (train.py:162-167, evaluated with the SAME D
    weights — the D update in between is the caller's business, see train_iteration).
    """
    out = {}
    g_forward, d_forward = FAMILY_FORWARDS[family]
    for p in PD.values():
        p.grad = None
    for p in PG.values():
        p.grad = None
    b = real.size(0)
    taps_real = {} if want_taps else None
    taps_real_in = {} if want_taps else None
    real_predict = d_forward(PD, real, step, alpha, taps=taps_real, taps_in=taps_real_in)
    real_loss = real_predict.mean() - 0.001 * (real_predict ** 2).mean()      # :128-129
    (-real_loss).backward()                                                    # .backward(mone) :130
    taps_g = {} if want_taps else None
    taps_g_in = {} if want_taps else None
    fake = g_forward(PG, z, step, alpha, tanh, pixelnorm, taps=taps_g, taps_in=taps_g_in)   # :135
    fake_predict = d_forward(PD, fake.detach(), step, alpha).mean()           # :136-138
    fake_predict.backward()                                                    # :139
    x_hat = (eps * real.data + (1 - eps) * fake.detach().data).requires_grad_(True)   # :142-144
    hat_predict = d_forward(PD, x_hat, step, alpha)
    grad_x_hat = torch.autograd.grad(outputs=hat_predict.sum(), inputs=x_hat,
                                     create_graph=True)[0]                     # :146-147
    gp = ((grad_x_hat.view(b, -1).norm(2, dim=1) - 1) ** 2).mean() * gp_lambda # :148-150
    gp.backward()                                                              # :151
    out['d_grads'] = {k: p.grad.clone() for k, p in PD.items() if p.grad is not None}
    out.update(real_predict=real_predict.detach(), fake=fake.detach(),
               fake_predict=fake_predict.detach(), x_hat=x_hat.detach(),
               hat_predict=hat_predict.detach(), grad_x_hat=grad_x_hat.detach(),
               grad_penalty=gp.detach(), disc_loss=(real_loss - fake_predict).detach())
    if want_taps:
        out['taps_d_real'] = {k: v.detach() for k, v in taps_real.items()}
        out['taps_g'] = {k: v.detach() for k, v in taps_g.items()}
        out['taps_d_real_in'] = {k: v.detach() for k, v in taps_real_in.items()}
        out['taps_g_in'] = {k: v.detach() for k, v in taps_g_in.items()}
    return out, fake


def g_phase(PG, PD, fake, step, alpha, family='base'):
    """train.py:158-167: loss = -D(fake).mean() through the stored G graph."""
    for p in list(PG.values()) + list(PD.values()):
        p.grad = None
    predict = FAMILY_FORWARDS[family][1](PD, fake, step, alpha)
    loss = -predict.mean()
    loss.backward()
    return loss.detach(), {k: p.grad.clone() for k, p in PG.items() if p.grad is not None}


class AdamState:
    """torch.optim.Adam(lr, betas=(0.0, 0.99)) restated (train.py:256-257); params with
    grad None are skipped, exactly as torch's optimiser does."""

    def __init__(self, P, lr=1e-3, betas=(0.0, 0.99), eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.t = {k: 0 for k in P}
        self.m = {k: torch.zeros_like(p) for k, p in P.items()}
        self.v = {k: torch.zeros_like(p) for k, p in P.items()}

    @torch.no_grad()
    def step(self, P, grads):
        for k, g in grads.items():
            self.t[k] += 1
            t = self.t[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1, bc2 = 1 - self.b1 ** t, 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / sqrt(bc2)).add_(self.eps)
            P[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)


@torch.no_grad()
def ema_accumulate(P_running, P_src, decay=0.999):
    """accumulate(), train.py:17-22."""
    for k in P_running:
        P_running[k].mul_(decay).add_(P_src[k], alpha=1 - decay)


def train_iteration(PG, PD, PG_run, optG, optD, real, z, eps, step, alpha, tanh=False,
                    pixelnorm=True, family='base'):
    """Full iteration train.py:122-169 (n_critic = 1): D phase, D Adam, G phase with the
    updated D, G Adam, EMA."""
    res, fake = train_step(PG, PD, real, z, eps, step, alpha, tanh, pixelnorm, family=family)
    optD.step(PD, res['d_grads'])
    gen_loss, g_grads = g_phase(PG, PD, fake, step, alpha, family)
    optG.step(PG, g_grads)
    ema_accumulate(PG_run, PG)
    res['gen_loss'] = gen_loss
    res['g_grads'] = g_grads
    return res
