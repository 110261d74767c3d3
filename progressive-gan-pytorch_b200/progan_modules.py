"""Host-side mirror of the reference's progan_modules.py for the training hot path.

Same class names, constructor arguments, attribute names, forward signatures
(`Generator(...).forward(input, step=0, alpha=-1)`, progan_modules.py:172,219;
`Discriminator(...).forward(input, step=0, alpha=-1)`, :258,282) and state-dict keys
(`...conv.weight_orig`, `...conv.bias`, `linear.linear.weight_orig`; SURVEY.md Appx A),
so the reference's train scripts and checkpoints work unchanged.  Everything below the
signatures is different: activations live in NHWC (bf16 in product mode, fp32 in check
mode), the equalized-LR scale sqrt(2/fan_in) (EqualLR.compute_weight, :22-27) is folded
into kernel epilogues instead of a per-forward weight multiply, and every op is an
explicit sm_100a kernel behind the C-ABI (see functions.py / include/progan_b200.h).
"""
from math import sqrt

import torch
from torch import nn

from . import functions as F_
from .kernels import ConvOp

_DEFAULT_PRECISION = "bf16"


def set_default_precision(p):
    """'bf16' (product: bf16 activations, fp32 accumulate) or 'fp32' (check mode)."""
    global _DEFAULT_PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _DEFAULT_PRECISION = p


def _act_dtype(p):
    # 'fp64' exists for the CPU test double only; the CUDA kernels take bf16 / fp32.
    return {"bf16": torch.bfloat16, "fp32": torch.float32, "fp64": torch.float64}[p]


def _img_dtype(p):
    return torch.float64 if p == "fp64" else torch.float32


class _Holder(nn.Module):
    """Holds weight_orig/bias under the name `.conv` / `.linear` like equal_lr() does."""

    def __init__(self, wshape, bias_n):
        super().__init__()
        # registration order (bias, weight_orig) matches the reference after equal_lr()
        self.bias = nn.Parameter(torch.zeros(bias_n))           # conv.bias.data.zero_()
        self.weight_orig = nn.Parameter(torch.randn(*wshape))   # conv.weight.data.normal_()


class EqualConv2d(nn.Module):
    """nn.Conv2d with equalized LR (progan_modules.py:63-73); kernel 1, 3 (pad 1) or 4 (pad 0)."""

    def __init__(self, in_channel, out_channel, kernel_size, padding=0):
        super().__init__()
        self.cin, self.cout, self.k, self.pad = in_channel, out_channel, kernel_size, padding
        self.conv = _Holder((out_channel, in_channel, kernel_size, kernel_size), out_channel)
        self.scale = sqrt(2 / (in_channel * kernel_size * kernel_size))
        self.op = ConvOp(kernel_size, padding, False, False)


class EqualConvTranspose2d(nn.Module):
    """nn.ConvTranspose2d(stride 1) with equalized LR (:76-92).  fan_in follows the
    reference quirk: weight is IOHW so size(1)*k*k = Cout*k*k (:24)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("only stride-1 ConvTranspose2d is on the hot path")
        self.cin, self.cout, self.k = in_channel, out_channel, kernel_size
        self.conv = _Holder((in_channel, out_channel, kernel_size, kernel_size), out_channel)
        self.scale = sqrt(2 / (out_channel * kernel_size * kernel_size))
        self.op = ConvOp(kernel_size, kernel_size - 1 - padding, True, True)


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear = _Holder((out_dim, in_dim), out_dim)
        self.in_dim, self.out_dim = in_dim, out_dim
        self.scale = sqrt(2 / in_dim)


class PixelNorm(nn.Module):
    """Marker only: PixelNorm (:54-60) is fused into the preceding conv's epilogue."""


class _LeakyMarker(nn.Module):
    def __init__(self, slope):
        super().__init__()
        self.negative_slope = slope


def _fused_layer(x, conv, slope, use_pn, pool=False, prev_link=None, make_link=False):
    h = conv.conv
    op = conv.op
    if x.shape[-1] != conv.cin:       # zero-padded input channels (after minibatch-stddev)
        op = ConvOp(op.k, op.pad, op.swap, op.flip, x.shape[-1], 0)
    return F_.conv_act(x, h.weight_orig, h.bias, op, conv.scale, slope, use_pn, pool, prev_link, make_link)


class ConvBlock(nn.Module):
    """conv -> [PixelNorm] -> LeakyReLU(0.2) -> conv -> [PixelNorm] -> LeakyReLU(0.2)
    (progan_modules.py:120-148); each conv+PN+LReLU is ONE fused kernel."""

    def __init__(self, in_channel, out_channel, kernel_size, padding, kernel_size2=None,
                 padding2=None, pixel_norm=True):
        super().__init__()
        pad2 = padding if padding2 is None else padding2
        k2 = kernel_size if kernel_size2 is None else kernel_size2
        mods = [EqualConv2d(in_channel, out_channel, kernel_size, padding=padding)]
        if pixel_norm:
            mods.append(PixelNorm())
        mods.append(_LeakyMarker(0.2))
        mods.append(EqualConv2d(out_channel, out_channel, k2, padding=pad2))
        if pixel_norm:
            mods.append(PixelNorm())
        mods.append(_LeakyMarker(0.2))
        self.conv = nn.Sequential(*mods)   # same indices as the reference: conv.0 / conv.3 (or conv.2)
        self.pixel_norm = pixel_norm
        self._c1 = 0
        self._c2 = 3 if pixel_norm else 2

    def forward(self, x, pool=False):
        """pool=True: followed by the x0.5 bilinear (= 2x2 average) downsample of the
        discriminator (progan_modules.py:299); its backward is fused with the activation's."""
        # the first activation has exactly one consumer (the second conv): its backward is fused
        # into that conv's data-gradient kernel (functions.ActLink)
        x, link = _fused_layer(x, self.conv[self._c1], 0.2, self.pixel_norm, make_link=True)
        return _fused_layer(x, self.conv[self._c2], 0.2, self.pixel_norm, pool, prev_link=link)


class MnistConvBlock(nn.Module):
    """conv -> [PixelNorm] -> LeakyReLU(0.2): the single-conv block of the mnist models
    (progan_modules.py:151-164); state-dict index conv.0."""

    def __init__(self, in_channel, out_channel, kernel_size, padding, pixel_norm=True):
        super().__init__()
        mods = [EqualConv2d(in_channel, out_channel, kernel_size, padding=padding)]
        if pixel_norm:
            mods.append(PixelNorm())
        mods.append(_LeakyMarker(0.2))
        self.conv = nn.Sequential(*mods)
        self.pixel_norm = pixel_norm

    def forward(self, x, pool=False):
        return _fused_layer(x, self.conv[0], 0.2, self.pixel_norm, pool)


def _fading(alpha):
    """(fading, alpha): the reference's gate `0 <= alpha < 1` (progan_modules.py:210,301).  A
    tensor alpha is evaluated like a number (the reference's comparison also works for 0-dim
    tensors) unless it is the Trainer's device scalar (marked `_pg_fading`), which stands for
    "the fade is active, read the value on the device" so that captured graphs follow the schedule."""
    if torch.is_tensor(alpha):
        if getattr(alpha, "_pg_fading", False):
            return True, alpha
        alpha = float(alpha)
    return bool(0 <= alpha < 1), alpha


class _AlphaMixin:
    @staticmethod
    def _alpha(alpha, device):
        """alpha lives in device memory so blend kernels (and captured graphs) read the
        current value; `alpha` may be a python float or an fp32 device tensor."""
        if torch.is_tensor(alpha):
            return alpha
        return torch.full((), float(alpha), device=device, dtype=torch.float32)


def _to_rgb(feat, m, act_dtype):
    h = m.conv
    C = feat.shape[-1]
    return F_.PwFwd.apply(feat, h.weight_orig, h.bias, "reduce", C, m.cout, 1, C, m.scale, act_dtype)


def _from_rgb(img, m, act_dtype):
    h = m.conv
    Kc = img.shape[1]
    return F_.PwFwd.apply(img, h.weight_orig, h.bias, "expand", m.cout, Kc, Kc, 1, m.scale, act_dtype)


class Generator(nn.Module, _AlphaMixin):
    def __init__(self, input_code_dim=128, in_channel=128, pixel_norm=True, tanh=True, max_step=6,
                 precision=None):
        super().__init__()
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.pixel_norm = pixel_norm
        self.precision = precision or _DEFAULT_PRECISION
        c = in_channel
        self.input_layer = nn.Sequential(EqualConvTranspose2d(input_code_dim, c, 4, 1, 0),
                                         PixelNorm(), _LeakyMarker(0.2))
        self.progression_4 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_8 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_64 = ConvBlock(c, c // 2, 3, 1, pixel_norm=pixel_norm)
        self.progression_128 = ConvBlock(c // 2, c // 4, 3, 1, pixel_norm=pixel_norm)
        self.progression_256 = ConvBlock(c // 4, c // 4, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_8 = EqualConv2d(c, 3, 1)
        self.to_rgb_16 = EqualConv2d(c, 3, 1)
        self.to_rgb_32 = EqualConv2d(c, 3, 1)
        self.to_rgb_64 = EqualConv2d(c // 2, 3, 1)
        self.to_rgb_128 = EqualConv2d(c // 4, 3, 1)
        self.to_rgb_256 = EqualConv2d(c // 4, 3, 1)
        self.max_step = max_step

    def _blocks(self):
        return [self.progression_8, self.progression_16, self.progression_32,
                self.progression_64, self.progression_128, self.progression_256]

    def _heads(self):
        return [self.to_rgb_8, self.to_rgb_16, self.to_rgb_32, self.to_rgb_64,
                self.to_rgb_128, self.to_rgb_256]

    def forward(self, input, step=0, alpha=-1):
        if step > self.max_step:
            step = self.max_step
        if step < 1:
            return None                      # reference falls through every `if` (:231-254)
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        z = input.reshape(-1, 1, 1, self.input_dim).to(dt).contiguous()
        # input layer always applies PixelNorm (:181-184), independent of `pixel_norm`
        feat = _fused_layer(z, self.input_layer[0], 0.2, True)
        feat = self.progression_4(feat)
        prev = None
        blocks, heads = self._blocks(), self._heads()
        for s in range(1, step + 1):
            prev = feat
            feat = blocks[s - 1](F_.upsample2(feat))
        out = _to_rgb(feat, heads[step - 1], dt)
        if step >= 2 and fading:
            skip = F_.upsample2(_to_rgb(prev, heads[step - 2], dt), "nchw")
            out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        if self.tanh:
            out = F_.Tanh.apply(out)
        return out


class Discriminator(nn.Module, _AlphaMixin):
    def __init__(self, feat_dim=128, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.precision = precision or _DEFAULT_PRECISION
        f = feat_dim
        self.progression = nn.ModuleList([ConvBlock(f // 4, f // 4, 3, 1),
                                          ConvBlock(f // 4, f // 2, 3, 1),
                                          ConvBlock(f // 2, f, 3, 1),
                                          ConvBlock(f, f, 3, 1),
                                          ConvBlock(f, f, 3, 1),
                                          ConvBlock(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.from_rgb = nn.ModuleList([EqualConv2d(3, f // 4, 1),
                                       EqualConv2d(3, f // 4, 1),
                                       EqualConv2d(3, f // 2, 1),
                                       EqualConv2d(3, f, 1),
                                       EqualConv2d(3, f, 1),
                                       EqualConv2d(3, f, 1),
                                       EqualConv2d(3, f, 1)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)
        # (block index, hook): the data-parallel Trainer learns through this tensor hook when a
        # backward sweep has passed every layer above that block (their gradients are final and
        # can be all-reduced while the sweep continues below)
        self._bwd_tap = None

    def forward(self, input, step=0, alpha=-1, mbstd_group=None):
        """mbstd_group (extension, default = reference behaviour): batch slices of this size get
        their own minibatch-stddev statistic, so one call on cat([real, fake]) equals the
        reference's two calls (train.py:126,137)."""
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        x = input.contiguous()
        if x.dtype != _img_dtype(self.precision):
            x = x.to(_img_dtype(self.precision))
        out = None
        for i in range(step, -1, -1):
            index = self.n_layer - i - 1
            if i == step:
                out = _from_rgb(x, self.from_rgb[index], dt)
            if i == 0:
                out = F_.Mbstd.apply(out, F_.K().mbstd_channels(out.shape[-1], out.dtype), mbstd_group)
            out = self.progression[index](out, pool=(i > 0))
            if i > 0:
                if i == step and fading:
                    skip = _from_rgb(F_.avgpool2(x, "nchw"), self.from_rgb[index + 1], dt)
                    out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
            if self._bwd_tap is not None and index == self._bwd_tap[0] and out.requires_grad:
                out.register_hook(self._bwd_tap[1])
        lin = self.linear.linear
        C = out.shape[-1]
        d = F_.PwFwd.apply(out, lin.weight_orig, lin.bias, "reduce", C, 1, 1, C, self.linear.scale, dt)
        return d.view(-1, 1)


class CorrectGenerator(nn.Module, _AlphaMixin):
    """The reference's CorrectGenerator (progan_modules.py:479-552; cifar_train.py /
    proper_cifar_train.py): 4x4 stem = ConvTranspose + conv in one Sequential, three ConvBlocks
    up to 32 px, a to_rgb head per resolution (4 px included), `step` counted from 1 = 4 px.
    Same kernels as Generator; same state-dict keys/shapes/order as the reference class."""

    def __init__(self, input_code_dim=512, in_channel=512, pixel_norm=True, tanh=False, max_step=4,
                 precision=None):
        super().__init__()
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.pixel_norm = pixel_norm
        self.precision = precision or _DEFAULT_PRECISION
        c = in_channel
        self.progression_4 = nn.Sequential(EqualConvTranspose2d(input_code_dim, c, 4, 1, 0), PixelNorm(),
                                           _LeakyMarker(0.2), EqualConv2d(c, c, 3, padding=1),
                                           PixelNorm(), _LeakyMarker(0.2))
        self.progression_8 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_4 = EqualConv2d(c, 3, 1)
        self.to_rgb_8 = EqualConv2d(c, 3, 1)
        self.to_rgb_16 = EqualConv2d(c, 3, 1)
        self.to_rgb_32 = EqualConv2d(c, 3, 1)
        self.max_step = max_step

    def _latent_dim(self):
        return self.input_dim

    def _output(self, feat1, feat2, head1, head2, alpha, fading, dt):
        """(:512-521): blend with the upsampled previous head iff 0 <= alpha < 1, then tanh."""
        out = _to_rgb(feat2, head2, dt)
        if fading:
            skip = F_.upsample2(_to_rgb(feat1, head1, dt), "nchw")
            out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        return F_.Tanh.apply(out) if self.tanh else out

    def forward(self, input, step=0, alpha=-1):
        if step > self.max_step:
            step = self.max_step
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        z = input.reshape(-1, 1, 1, self._latent_dim()).to(dt).contiguous()
        # the stem applies PixelNorm after both convs whatever `pixel_norm` says (:487-494)
        out_4 = _fused_layer(z, self.progression_4[0], 0.2, True)
        out_4 = _fused_layer(out_4, self.progression_4[3], 0.2, True)
        if step == 1:
            out = _to_rgb(out_4, self.to_rgb_4, dt)
            return F_.Tanh.apply(out) if self.tanh else out
        out_8 = self.progression_8(F_.upsample2(out_4))
        if step == 2:
            if self.tanh:                     # reference quirk (:534-537): no blend on this path
                return F_.Tanh.apply(_to_rgb(out_8, self.to_rgb_8, dt))
            return self._output(out_4, out_8, self.to_rgb_4, self.to_rgb_8, alpha, fading, dt)
        out_16 = self.progression_16(F_.upsample2(out_8))
        if step == 3:
            return self._output(out_8, out_16, self.to_rgb_8, self.to_rgb_16, alpha, fading, dt)
        out_32 = self.progression_32(F_.upsample2(out_16))
        if step == 4:
            return self._output(out_16, out_32, self.to_rgb_16, self.to_rgb_32, alpha, fading, dt)
        return None                           # step <= 0: the reference falls through every `if`


class CorrectDiscriminator(nn.Module, _AlphaMixin):
    """The reference's CorrectDiscriminator (progan_modules.py:555-598): four blocks of constant
    width, `step` counted from 1 = 4 px, fade-in at every step > 1, minibatch-stddev before the
    last block."""

    def __init__(self, feat_dim=512, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.precision = precision or _DEFAULT_PRECISION
        f = feat_dim
        self.progression = nn.ModuleList([ConvBlock(f, f, 3, 1), ConvBlock(f, f, 3, 1),
                                          ConvBlock(f, f, 3, 1), ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.from_rgb = nn.ModuleList([EqualConv2d(3, f, 1) for _ in range(4)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input, step=0, alpha=-1, mbstd_group=None):
        if step < 1:
            raise RuntimeError("CorrectDiscriminator: step must be >= 1 (the reference fails with an "
                               "unbound `out` for step 0, progan_modules.py:578-596)")
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        x = input.contiguous()
        if x.dtype != _img_dtype(self.precision):
            x = x.to(_img_dtype(self.precision))
        out = None
        for i in range(step, 0, -1):
            index = self.n_layer - i
            if i == step:
                out = _from_rgb(x, self.from_rgb[index], dt)
            if i == 1:
                out = F_.Mbstd.apply(out, F_.K().mbstd_channels(out.shape[-1], out.dtype), mbstd_group)
            out = self.progression[index](out, pool=(i > 1))
            if i > 1 and i == step and fading:
                skip = _from_rgb(F_.avgpool2(x, "nchw"), self.from_rgb[index + 1], dt)
                out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        lin = self.linear.linear
        C = out.shape[-1]
        d = F_.PwFwd.apply(out, lin.weight_orig, lin.bias, "reduce", C, 1, 1, C, self.linear.scale, dt)
        return d.view(-1, 1)


class _EmbedHolder(nn.Module):
    def __init__(self, n, d):
        super().__init__()
        self.weight_orig = nn.Parameter(torch.randn(n, d))       # embed.weight.data.normal_() (:113)


class EqualEmbed(nn.Module):
    """nn.Embedding with equalized LR (progan_modules.py:109-117): fan_in = embedding_dim (:24).
    The row gather is host-side glue (B x dim floats), not a kernel of the hot path."""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__()
        self.embed = _EmbedHolder(num_embeddings, embedding_dim)
        self.scale = sqrt(2 / embedding_dim)

    def forward(self, label):
        return torch.nn.functional.embedding(label, self.embed.weight_orig) * self.scale


class ConditionalCorrectGenerator(CorrectGenerator):
    """Class-conditional generator of conditional_proper_cifar_train.py / conditional_proper_wikiart.py
    (progan_modules.py:601-694): the label embedding (dimension = input_code_dim) is concatenated
    to z in front of the 4x4 ConvTranspose stem; six resolution steps (4 ... 128 px)."""

    def __init__(self, input_code_dim=512, num_of_classes=10, in_channel=512, pixel_norm=True, tanh=False,
                 max_step=4, do_equal_embed=False, precision=None):
        nn.Module.__init__(self)
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.pixel_norm = pixel_norm
        self.num_of_classes = num_of_classes
        self.embedding_dim = input_code_dim
        self.do_equal_embed = do_equal_embed
        self.precision = precision or _DEFAULT_PRECISION
        c = in_channel
        if do_equal_embed:
            self.embedding = EqualEmbed(num_of_classes, self.embedding_dim)
        else:
            self.embedding = nn.Embedding(num_of_classes, embedding_dim=self.embedding_dim)
        self.progression_4 = nn.Sequential(
            EqualConvTranspose2d(input_code_dim + self.embedding_dim, c, 4, 1, 0), PixelNorm(),
            _LeakyMarker(0.2), EqualConv2d(c, c, 3, padding=1), PixelNorm(), _LeakyMarker(0.2))
        self.progression_8 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_64 = ConvBlock(c, c // 2, 3, 1, pixel_norm=pixel_norm)
        self.progression_128 = ConvBlock(c // 2, c // 4, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_4 = EqualConv2d(c, 3, 1)
        self.to_rgb_8 = EqualConv2d(c, 3, 1)
        self.to_rgb_16 = EqualConv2d(c, 3, 1)
        self.to_rgb_32 = EqualConv2d(c, 3, 1)
        self.to_rgb_64 = EqualConv2d(c // 2, 3, 1)
        self.to_rgb_128 = EqualConv2d(c // 4, 3, 1)
        self.max_step = max_step

    def forward(self, input, label, step=0, alpha=-1):
        if step > self.max_step:
            step = self.max_step
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        data_in = torch.cat([input, self.embedding(label).to(input.dtype)], 1)        # (:664-668)
        z = data_in.reshape(-1, 1, 1, self.input_dim + self.embedding_dim).to(dt).contiguous()
        out_4 = _fused_layer(z, self.progression_4[0], 0.2, True)
        out_4 = _fused_layer(out_4, self.progression_4[3], 0.2, True)
        if step == 1:
            out = _to_rgb(out_4, self.to_rgb_4, dt)
            return F_.Tanh.apply(out) if self.tanh else out
        feats = [out_4]
        blocks = [self.progression_8, self.progression_16, self.progression_32, self.progression_64,
                  self.progression_128]
        heads = [self.to_rgb_4, self.to_rgb_8, self.to_rgb_16, self.to_rgb_32, self.to_rgb_64,
                 self.to_rgb_128]
        for s_ in range(2, 7):
            feats.append(blocks[s_ - 2](F_.upsample2(feats[-1])))
            if step == s_:
                if s_ == 2 and self.tanh:        # same quirk as CorrectGenerator (:674-676)
                    return F_.Tanh.apply(_to_rgb(feats[-1], self.to_rgb_8, dt))
                return self._output(feats[-2], feats[-1], heads[s_ - 2], heads[s_ - 1], alpha, fading, dt)
        return None


class ConditionalCorrectDiscriminatorWgangp(nn.Module, _AlphaMixin):
    """Class-conditional critic (progan_modules.py:697-775): a per-resolution label embedding of
    R*R values is appended to the image as a 4th channel in front of from_rgb (and of the fade-in
    skip from_rgb); six blocks, step 1 = 4 px ... 6 = 128 px."""

    def __init__(self, feat_dim=128, num_of_classes=10, do_equal_embed=False, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.num_of_classes = num_of_classes
        self.do_equal_embed = do_equal_embed
        self.precision = precision or _DEFAULT_PRECISION
        f = feat_dim
        self.progression = nn.ModuleList([ConvBlock(f // 4, f // 2, 3, 1), ConvBlock(f // 2, f, 3, 1),
                                          ConvBlock(f, f, 3, 1), ConvBlock(f, f, 3, 1), ConvBlock(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        sizes = [128 ** 2, 64 ** 2, 32 ** 2, 16 ** 2, 8 ** 2, 4 ** 2]
        if do_equal_embed:
            self.embeddings = nn.ModuleList([EqualEmbed(num_of_classes, d) for d in sizes])
        else:
            self.embeddings = nn.ModuleList([nn.Embedding(num_of_classes, d) for d in sizes])
        self.from_rgb = nn.ModuleList([EqualConv2d(3 + 1, c_, 1) for c_ in (f // 4, f // 2, f, f, f, f)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def _with_label_plane(self, img, label, index):
        plane = self.embeddings[index](label).to(img.dtype).view(-1, 1, img.shape[-2], img.shape[-1])
        return torch.cat([img, plane], 1).contiguous()                                # (:747-749)

    def forward(self, input, label, step=0, alpha=-1, mbstd_group=None):
        if step < 1:
            raise RuntimeError("ConditionalCorrectDiscriminatorWgangp: step must be >= 1")
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        x = input.contiguous()
        if x.dtype != _img_dtype(self.precision):
            x = x.to(_img_dtype(self.precision))
        out = None
        for i in range(step, 0, -1):
            index = self.n_layer - i
            if i == step:
                out = _from_rgb(self._with_label_plane(x, label, index), self.from_rgb[index], dt)
            if i == 1:
                out = F_.Mbstd.apply(out, F_.K().mbstd_channels(out.shape[-1], out.dtype), mbstd_group)
            out = self.progression[index](out, pool=(i > 1))
            if i > 1 and i == step and fading:
                skip_img = self._with_label_plane(F_.avgpool2(x, "nchw"), label, index + 1)   # (:766-768)
                skip = _from_rgb(skip_img, self.from_rgb[index + 1], dt)
                out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        lin = self.linear.linear
        C = out.shape[-1]
        d = F_.PwFwd.apply(out, lin.weight_orig, lin.bias, "reduce", C, 1, 1, C, self.linear.scale, dt)
        return d.view(-1, 1)


def _critic_trunk(m, input, step, alpha, mbstd_group, last, plane=None):
    """Body shared by the remaining critics: from_rgb at resolution `step`, the blocks down to the
    4x4 one (reached at i == `last`: 0 for the wiring of Discriminator, 1 for the Correct* wiring),
    fade-in blend with the from_rgb of the halved input at the first block, minibatch-stddev in
    front of the last block.  `plane(img, index)` appends the label plane where the class has
    one.  Returns the [B,1,1,C] feature in front of the linear head."""
    dt = _act_dtype(m.precision)
    fading, alpha = _fading(alpha)
    x = input.contiguous()
    if x.dtype != _img_dtype(m.precision):
        x = x.to(_img_dtype(m.precision))
    out = None
    for i in range(step, last - 1, -1):
        index = m.n_layer - i - 1 + last
        if i == step:
            out = _from_rgb(plane(x, index) if plane else x, m.from_rgb[index], dt)
        if i == last:
            out = F_.Mbstd.apply(out, F_.K().mbstd_channels(out.shape[-1], out.dtype), mbstd_group)
        out = m.progression[index](out, pool=(i > last))
        if i > last and i == step and fading:
            half = F_.avgpool2(x, "nchw")
            skip = _from_rgb(plane(half, index + 1) if plane else half, m.from_rgb[index + 1], dt)
            out = F_.Blend.apply(skip, out, m._alpha(alpha, out.device))
    return out, dt


def _critic_head(m, out, dt):
    lin = m.linear.linear
    C = out.shape[-1]
    return F_.PwFwd.apply(out, lin.weight_orig, lin.bias, "reduce", C, 1, 1, C, m.linear.scale, dt)


def _projection_score(m, out, label, dt):
    """Projection critic (progan_modules.py:909-913, mnist_pggan.py:338-343):
    D(x, y) = linear(h) + <h, normalize(embedding[y])>, returned with shape [B]."""
    embed = torch.nn.functional.normalize(m.embedding(label))
    proj = (out.reshape(out.shape[0], -1).to(embed.dtype) * embed).sum(dim=-1)
    return _critic_head(m, out, dt).view(-1) + proj.to(torch.float32)


def _label_plane(m, img, label, index):
    plane = m.embeddings[index](label).to(img.dtype).view(-1, 1, img.shape[-2], img.shape[-1])
    return torch.cat([img, plane], 1).contiguous()


class ConditionalGenerator(Generator):
    """Class-conditional form of Generator (progan_modules.py:314-404; conditional_cifar10_wgan_train.py):
    a plain nn.Embedding of dimension num_of_classes is concatenated to z in front of the input
    layer; everything behind it is Generator's wiring (step 1 = 8 px ... 6 = 256 px)."""

    def __init__(self, input_code_dim=128, num_of_classes=10, in_channel=128, pixel_norm=True, tanh=True,
                 max_step=6, precision=None):
        nn.Module.__init__(self)
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.pixel_norm = pixel_norm
        self.num_of_classes = num_of_classes
        self.embedding_dim = num_of_classes
        self.precision = precision or _DEFAULT_PRECISION
        c = in_channel
        self.embedding = nn.Embedding(num_of_classes, self.embedding_dim)
        self.input_layer = nn.Sequential(EqualConvTranspose2d(input_code_dim + self.embedding_dim, c, 4, 1, 0),
                                         PixelNorm(), _LeakyMarker(0.2))
        self.progression_4 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_8 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_64 = ConvBlock(c, c // 2, 3, 1, pixel_norm=pixel_norm)
        self.progression_128 = ConvBlock(c // 2, c // 4, 3, 1, pixel_norm=pixel_norm)
        self.progression_256 = ConvBlock(c // 4, c // 4, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_8 = EqualConv2d(c, 3, 1)
        self.to_rgb_16 = EqualConv2d(c, 3, 1)
        self.to_rgb_32 = EqualConv2d(c, 3, 1)
        self.to_rgb_64 = EqualConv2d(c // 2, 3, 1)
        self.to_rgb_128 = EqualConv2d(c // 4, 3, 1)
        self.to_rgb_256 = EqualConv2d(c // 4, 3, 1)
        self.max_step = max_step

    def _latent(self, input, label):
        return torch.cat([input, self.embedding(label).to(input.dtype)], 1)               # (:370-371)

    def forward(self, input, label, step=0, alpha=-1):
        if step > self.max_step:
            step = self.max_step
        if step < 1:
            return None
        dt = _act_dtype(self.precision)
        fading, alpha = _fading(alpha)
        z = self._latent(input, label).reshape(-1, 1, 1, self.input_dim + self.embedding_dim).to(dt).contiguous()
        feat = self.progression_4(_fused_layer(z, self.input_layer[0], 0.2, True))
        prev = None
        blocks, heads = self._blocks(), self._heads()
        for s in range(1, step + 1):
            prev = feat
            feat = blocks[s - 1](F_.upsample2(feat))
        out = _to_rgb(feat, heads[step - 1], dt)
        if step >= 2 and fading:
            skip = F_.upsample2(_to_rgb(prev, heads[step - 2], dt), "nchw")
            out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        return F_.Tanh.apply(out) if self.tanh else out


class ConditionalDiscriminatorWgangp(nn.Module, _AlphaMixin):
    """Class-conditional form of Discriminator (progan_modules.py:407-476): a per-resolution
    nn.Embedding of R*R values is appended to the image as a 4th channel in front of from_rgb (and
    of the fade-in skip from_rgb); seven blocks, step 0 = 4 px ... 6 = 256 px."""

    def __init__(self, feat_dim=128, num_of_classes=10, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.num_of_classes = num_of_classes
        self.precision = precision or _DEFAULT_PRECISION
        f = feat_dim
        self.progression = nn.ModuleList([ConvBlock(f // 4, f // 4, 3, 1), ConvBlock(f // 4, f // 2, 3, 1),
                                          ConvBlock(f // 2, f, 3, 1), ConvBlock(f, f, 3, 1),
                                          ConvBlock(f, f, 3, 1), ConvBlock(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.embeddings = nn.ModuleList([nn.Embedding(num_of_classes, r * r)
                                         for r in (256, 128, 64, 32, 16, 8, 4)])
        self.from_rgb = nn.ModuleList([EqualConv2d(3 + 1, c_, 1) for c_ in (f // 4, f // 4, f // 2, f, f, f, f)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input, label, step=0, alpha=-1, mbstd_group=None):
        out, dt = _critic_trunk(self, input, step, alpha, mbstd_group, 0,
                                lambda img, index: _label_plane(self, img, label, index))
        return _critic_head(self, out, dt).view(-1, 1)


class ConditionalCorrectGeneratorAda(CorrectGenerator):
    """ADA-style conditional generator (progan_modules.py:778-854): cat(normalize(z),
    normalize(embedding[y])) in front of CorrectGenerator's wiring (step 1 = 4 px ... 4 = 32 px)."""

    def __init__(self, input_code_dim=512, num_of_classes=10, in_channel=512, pixel_norm=True, tanh=False,
                 max_step=4, precision=None):
        nn.Module.__init__(self)
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.pixel_norm = pixel_norm
        self.num_of_classes = num_of_classes
        self.embedding_dim = input_code_dim
        self.precision = precision or _DEFAULT_PRECISION
        c = in_channel
        self.embedding = nn.Embedding(num_of_classes, embedding_dim=self.embedding_dim)
        self.progression_4 = nn.Sequential(
            EqualConvTranspose2d(input_code_dim + self.embedding_dim, c, 4, 1, 0), PixelNorm(),
            _LeakyMarker(0.2), EqualConv2d(c, c, 3, padding=1), PixelNorm(), _LeakyMarker(0.2))
        self.progression_8 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = ConvBlock(c, c, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_4 = EqualConv2d(c, 3, 1)
        self.to_rgb_8 = EqualConv2d(c, 3, 1)
        self.to_rgb_16 = EqualConv2d(c, 3, 1)
        self.to_rgb_32 = EqualConv2d(c, 3, 1)
        self.max_step = max_step

    def _latent_dim(self):
        return self.input_dim + self.embedding_dim

    def forward(self, input, label, step=0, alpha=-1):
        nrm = torch.nn.functional.normalize
        data_in = torch.cat([nrm(input), nrm(self.embedding(label)).to(input.dtype)], 1)  # (:830-833)
        return CorrectGenerator.forward(self, data_in, step, alpha)


class ConditionalCorrectDiscriminatorAda(nn.Module, _AlphaMixin):
    """Projection critic on CorrectDiscriminator's wiring (progan_modules.py:857-915): plain
    3-channel from_rgb, the label enters only through <h, normalize(embedding[y])> added to the
    linear score; returns shape [B]."""

    def __init__(self, feat_dim=512, num_of_classes=10, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.num_of_classes = num_of_classes
        self.embedding_dim = feat_dim
        self.precision = precision or _DEFAULT_PRECISION
        f = feat_dim
        self.embedding = nn.Embedding(num_of_classes, embedding_dim=self.embedding_dim)
        self.progression = nn.ModuleList([ConvBlock(f, f, 3, 1), ConvBlock(f, f, 3, 1),
                                          ConvBlock(f, f, 3, 1), ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.from_rgb = nn.ModuleList([EqualConv2d(3, f, 1) for _ in range(4)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input, label, step=0, alpha=-1, mbstd_group=None):
        if step < 1:
            raise RuntimeError("ConditionalCorrectDiscriminatorAda: step must be >= 1")
        out, dt = _critic_trunk(self, input, step, alpha, mbstd_group, 1)
        return _projection_score(self, out, label, dt)
