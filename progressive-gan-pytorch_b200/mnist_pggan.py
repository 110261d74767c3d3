"""Host-side mirror of the reference's mnist_pggan.py (BASELINE config 0: 1-channel images,
progressive growing 8 -> 32 px): `Generator(input_code_dim=128, in_channel=64, pixel_norm=True,
tanh=True, use_mnist_conv_blocks=True)` (mnist_pggan.py:10-80) and `Discriminator(feat_dim=64,
use_mnist_conv_blocks=True)` (:83-137).  Same kernels as progan_modules; same constructor
arguments, attributes, forward signatures and state-dict keys/shapes/order as the reference —
including its quirks: LeakyReLU(0.1) after the input layer (:21), `max_step = 3` (:35), and the
two dead `mnist_progression_*` blocks the discriminator carries for checkpoint compatibility
(:93-97) — plus the class-conditional variants of conditional_mnist_wgan_train.py (:140-345)."""
import torch
from torch import nn

from . import functions as F_
from .progan_modules import (ConvBlock, EqualConv2d, EqualConvTranspose2d, EqualLinear, MnistConvBlock,
                             PixelNorm, _AlphaMixin, _DEFAULT_PRECISION, _LeakyMarker, _act_dtype,
                             _fading, _from_rgb, _fused_layer, _img_dtype, _to_rgb)
from . import progan_modules as _pm


class Generator(nn.Module, _AlphaMixin):
    def __init__(self, input_code_dim=128, in_channel=64, pixel_norm=True, tanh=True,
                 use_mnist_conv_blocks=True, precision=None):
        super().__init__()
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.use_mnist_conv_blocks = use_mnist_conv_blocks
        self.pixel_norm = pixel_norm
        self.precision = precision or _pm._DEFAULT_PRECISION
        c = in_channel
        self.input_layer = nn.Sequential(EqualConvTranspose2d(input_code_dim, c, 4, 1, 0), PixelNorm(),
                                         _LeakyMarker(0.1))
        block = MnistConvBlock if use_mnist_conv_blocks else ConvBlock
        self.progression_4 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_8 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_8 = EqualConv2d(c, 1, 1)
        self.to_rgb_16 = EqualConv2d(c, 1, 1)
        self.to_rgb_32 = EqualConv2d(c, 1, 1)
        self.max_step = 3

    def _latent_dim(self):
        return self.input_dim

    def _output(self, feat1, feat2, head1, head2, alpha, fading, dt):
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
        out_4 = _fused_layer(z, self.input_layer[0], 0.1, True)          # slope 0.1 (:21)
        out_4 = self.progression_4(out_4)
        out_8 = self.progression_8(F_.upsample2(out_4))
        if step == 1:
            out = _to_rgb(out_8, self.to_rgb_8, dt)
            return F_.Tanh.apply(out) if self.tanh else out
        out_16 = self.progression_16(F_.upsample2(out_8))
        if step == 2:
            return self._output(out_8, out_16, self.to_rgb_8, self.to_rgb_16, alpha, fading, dt)
        out_32 = self.progression_32(F_.upsample2(out_16))
        if step == 3:
            return self._output(out_16, out_32, self.to_rgb_16, self.to_rgb_32, alpha, fading, dt)
        return None            # step <= 0 (the reference's step-4 branch is unreachable: max_step = 3)


class Discriminator(nn.Module, _AlphaMixin):
    def __init__(self, feat_dim=64, use_mnist_conv_blocks=True, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.use_mnist_conv_blocks = use_mnist_conv_blocks
        self.precision = precision or _pm._DEFAULT_PRECISION
        f = feat_dim
        block = MnistConvBlock if use_mnist_conv_blocks else ConvBlock
        self.progression = nn.ModuleList([block(f, f, 3, 1), block(f, f, 3, 1), block(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        # dead blocks kept by the reference for old checkpoints (:93-97): parameters only
        self.mnist_progression_0 = MnistConvBlock(f + 1, f, 3, 1)
        self.mnist_progression_1 = MnistConvBlock(f + 1, f, 4, 0)
        self.from_rgb = nn.ModuleList([EqualConv2d(1, f, 1) for _ in range(4)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input, step=0, alpha=-1, mbstd_group=None):
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
            if i > 0 and i == step and fading:
                skip = _from_rgb(F_.avgpool2(x, "nchw"), self.from_rgb[index + 1], dt)
                out = F_.Blend.apply(skip, out, self._alpha(alpha, out.device))
        lin = self.linear.linear
        C = out.shape[-1]
        d = F_.PwFwd.apply(out, lin.weight_orig, lin.bias, "reduce", C, 1, 1, C, self.linear.scale, dt)
        return d.view(-1, 1)


class ConditionalGenerator(Generator):
    """mnist_pggan.ConditionalGenerator (mnist_pggan.py:140-221; conditional_mnist_wgan_train.py):
    cat(normalize(z), normalize(embedding[y])) in front of Generator's wiring; embedding_dim =
    input_code_dim."""

    def __init__(self, input_code_dim=128, num_of_classes=10, in_channel=64, pixel_norm=True, tanh=True,
                 use_mnist_conv_blocks=True, precision=None):
        nn.Module.__init__(self)
        self.input_dim = input_code_dim
        self.in_channel = in_channel
        self.tanh = tanh
        self.use_mnist_conv_blocks = use_mnist_conv_blocks
        self.num_of_classes = num_of_classes
        self.embedding_dim = input_code_dim
        self.pixel_norm = pixel_norm
        self.precision = precision or _pm._DEFAULT_PRECISION
        c = in_channel
        self.embedding = nn.Embedding(num_of_classes, self.embedding_dim)
        self.input_layer = nn.Sequential(EqualConvTranspose2d(input_code_dim + self.embedding_dim, c, 4, 1, 0),
                                         PixelNorm(), _LeakyMarker(0.1))
        block = MnistConvBlock if use_mnist_conv_blocks else ConvBlock
        self.progression_4 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_8 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_16 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.progression_32 = block(c, c, 3, 1, pixel_norm=pixel_norm)
        self.to_rgb_8 = EqualConv2d(c, 1, 1)
        self.to_rgb_16 = EqualConv2d(c, 1, 1)
        self.to_rgb_32 = EqualConv2d(c, 1, 1)
        self.max_step = 3

    def _latent_dim(self):
        return self.input_dim + self.embedding_dim

    def forward(self, input, label, step=0, alpha=-1):
        nrm = torch.nn.functional.normalize
        data_in = torch.cat([nrm(input), nrm(self.embedding(label)).to(input.dtype)], 1)   # (:192-196)
        return Generator.forward(self, data_in, step, alpha)


class ConditionalDiscriminatorWgangp(nn.Module, _AlphaMixin):
    """mnist_pggan.ConditionalDiscriminatorWgangp (mnist_pggan.py:224-286): the label enters as a
    second image channel (an R*R embedding per resolution) in front of every from_rgb."""

    def __init__(self, feat_dim=64, num_of_classes=10, use_mnist_conv_blocks=True, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.num_of_classes = num_of_classes
        self.use_mnist_conv_blocks = use_mnist_conv_blocks
        self.precision = precision or _pm._DEFAULT_PRECISION
        f = feat_dim
        block = MnistConvBlock if use_mnist_conv_blocks else ConvBlock
        self.progression = nn.ModuleList([block(f, f, 3, 1), block(f, f, 3, 1), block(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.embeddings = nn.ModuleList([nn.Embedding(num_of_classes, r * r) for r in (32, 16, 8, 4)])
        self.from_rgb = nn.ModuleList([EqualConv2d(1 + 1, f, 1) for _ in range(4)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input_data, label, step=0, alpha=-1, mbstd_group=None):
        out, dt = _pm._critic_trunk(self, input_data, step, alpha, mbstd_group, 0,
                                    lambda img, index: _pm._label_plane(self, img, label, index))
        return _pm._critic_head(self, out, dt).view(-1, 1)


class ConditionalDiscriminatorAda(nn.Module, _AlphaMixin):
    """mnist_pggan.ConditionalDiscriminatorAda (mnist_pggan.py:289-345): projection critic, the
    label enters through <h, normalize(embedding[y])>; returns shape [B]."""

    def __init__(self, feat_dim=64, num_of_classes=10, use_mnist_conv_blocks=True, precision=None):
        super().__init__()
        self.feat_dim = feat_dim
        self.num_of_classes = num_of_classes
        self.embedding_dim = feat_dim
        self.use_mnist_conv_blocks = use_mnist_conv_blocks
        self.precision = precision or _pm._DEFAULT_PRECISION
        f = feat_dim
        block = MnistConvBlock if use_mnist_conv_blocks else ConvBlock
        self.progression = nn.ModuleList([block(f, f, 3, 1), block(f, f, 3, 1), block(f, f, 3, 1),
                                          ConvBlock(f + 1, f, 3, 1, 4, 0)])
        self.embedding = nn.Embedding(num_of_classes, embedding_dim=self.embedding_dim)
        self.from_rgb = nn.ModuleList([EqualConv2d(1, f, 1) for _ in range(4)])
        self.n_layer = len(self.progression)
        self.linear = EqualLinear(f, 1)

    def forward(self, input_data, label, step=0, alpha=-1, mbstd_group=None):
        out, dt = _pm._critic_trunk(self, input_data, step, alpha, mbstd_group, 0)
        return _pm._projection_score(self, out, label, dt)
