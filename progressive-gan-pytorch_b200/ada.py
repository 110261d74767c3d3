"""ADA augmentation between the images and the critic.

Mirror of the reference's vendored StyleGAN2-ADA pipeline — `ada/augment.py:118-428`
(`AugmentPipe`: pixel blitting, general geometric transforms with sym6 anti-aliasing, colour
transforms, sym2 band filters, noise, cutout) and `ada/adapt_augm.py:6-51` (`AdaptiveAugment`, the
ADA-p controller) — with the same constructor arguments, attribute and buffer names (`p`,
`Hz_geom`, `Hz_fbank`), the same random-number consumption (same draws in the same order, so a
seeded run reproduces the reference's output) and the same arithmetic.

What is different:
  * parameter SAMPLING (`AugmentPipe.sample`) and EXECUTION (`AugmentPipe.apply`) are separate:
    the sampled per-image transforms (3x3 geometry, 4x4 colour, band gains, noise field, cutout
    boxes) are a plain object a caller can keep, replay on the fake / x_hat batches, or inject;
  * the pipe is **twice differentiable** w.r.t. the images, which the WGAN-GP path needs when the
    augmentation sits between x_hat and D (train.py:142-151).  Every stage is linear in the images;
    the bilinear resampling is an explicit operator pair (`_Resample` / `_ResampleT`, each the
    other's backward) because `aten::grid_sampler_2d_backward` has no derivative in torch 2.x and
    the reference's own `grid_sample_gradfix` is only active on torch 1.7-1.9
    (`ada/torch_utils/ops/grid_sample_gradfix.py:37`): on this torch the reference pipe fails in
    the second backward;
  * the FIR resampling steps (`upfirdn2d.upsample2d / downsample2d`, `ops/upfirdn2d.py:300-384`)
    are one separable helper on depthwise convolutions.

Device path: ATen ops on the images' device (no hand-written kernels: this component was written
after the round's GPU budget was spent; it is checked on the CPU against the live reference and
against committed golden vectors, `tests/test_ada.py`).
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

# Low-pass decomposition filters of the symlet wavelets used by the reference
# (ada/augment.py:22-40: 'sym6' anti-aliases the geometric warp, 'sym2' builds the band filters).
_SYM2 = [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025]
_SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
         0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
         -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]


# ------------------------------------------------------------------ small batched matrices
def _mat(rows, like=None, device=None):
    """Matrix from python rows whose entries are numbers or batch tensors of one shape
    ([...]-shaped entries give a [..., R, C] result); plain numbers only: a constant matrix."""
    flat = [e for r in rows for e in r]
    tens = [e for e in flat if torch.is_tensor(e)]
    if not tens:
        return torch.tensor(rows, dtype=torch.float32, device=device)
    ref = tens[0]
    cols = [e if torch.is_tensor(e) else torch.full(ref.shape, float(e), dtype=torch.float32, device=ref.device)
            for e in flat]
    return torch.stack(cols, dim=-1).reshape(ref.shape + (len(rows), len(rows[0])))


def _shift2(tx, ty, **kw):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], **kw)


def _scale2(sx, sy, **kw):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], **kw)


def _rot2(theta, **kw):
    return _mat([[torch.cos(theta), torch.sin(-theta), 0], [torch.sin(theta), torch.cos(theta), 0], [0, 0, 1]], **kw)


def _shift3(tx, ty, tz, **kw):
    return _mat([[1, 0, 0, tx], [0, 1, 0, ty], [0, 0, 1, tz], [0, 0, 0, 1]], **kw)


def _scale3(sx, sy, sz, **kw):
    return _mat([[sx, 0, 0, 0], [0, sy, 0, 0], [0, 0, sz, 0], [0, 0, 0, 1]], **kw)


def _rot3(axis, theta, **kw):
    """Rotation by theta about the unit axis (first three entries of `axis`), homogeneous 4x4."""
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    s, c = torch.sin(theta), torch.cos(theta)
    k = 1 - c
    return _mat([[x * x * k + c, x * y * k - z * s, x * z * k + y * s, 0],
                 [y * x * k + z * s, y * y * k + c, y * z * k - x * s, 0],
                 [z * x * k - y * s, z * y * k + x * s, z * z * k + c, 0],
                 [0, 0, 0, 1]], **kw)


# ------------------------------------------------------------------ linear resampling operators
class _Resample(torch.autograd.Function):
    """y = L_grid x: bilinear gather with zero padding (F.grid_sample, align_corners=False).
    Linear in x; its backward is the transposed operator, whose backward is this one again — so
    any number of derivatives w.r.t. the images exists."""

    @staticmethod
    def forward(ctx, x, grid):
        ctx.save_for_backward(grid)
        ctx.in_shape = x.shape
        return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)

    @staticmethod
    def backward(ctx, gy):
        (grid,) = ctx.saved_tensors
        return _ResampleT.apply(gy, grid, ctx.in_shape), None


class _ResampleT(torch.autograd.Function):
    """gx = L_grid^T gy: the bilinear scatter (what aten::grid_sampler_2d_backward computes for
    the input)."""

    @staticmethod
    def forward(ctx, gy, grid, in_shape):
        ctx.save_for_backward(grid)
        probe = gy.new_zeros(in_shape)                   # only its shape / dtype are used
        gx, _ = torch.ops.aten.grid_sampler_2d_backward(gy.contiguous(), probe, grid, 0, 0, False, (True, False))
        return gx

    @staticmethod
    def backward(ctx, ggx):
        (grid,) = ctx.saved_tensors
        return _Resample.apply(ggx, grid), None, None


def _fir_axis_taps(f, flip):
    """The 1-D taps as F.conv2d (a cross-correlation) must see them: a true convolution unless
    `flip` (upfirdn2d's flip_filter: False = convolution, True = correlation)."""
    return f if flip else f.flip(0)


def fir_resample(x, f, up=1, down=1, pad=(0, 0, 0, 0), flip=False, gain=1.0):
    """Zero-insert upsampling by `up`, zero padding / cropping (x0, x1, y0, y1; negative = crop),
    separable FIR filter `f` (1-D taps, already normalised), decimation by `down`
    (ada/torch_utils/ops/upfirdn2d.py:168-208, the separable branch)."""
    n, c, h, w = x.shape
    if up > 1:
        x = x.reshape(n, c, h, 1, w, 1)
        x = F.pad(x, [0, up - 1, 0, 0, 0, up - 1])
        x = x.reshape(n, c, h * up, w * up)
    x0, x1, y0, y1 = pad
    x = F.pad(x, [max(x0, 0), max(x1, 0), max(y0, 0), max(y1, 0)])
    x = x[:, :, max(-y0, 0): x.shape[2] - max(-y1, 0), max(-x0, 0): x.shape[3] - max(-x1, 0)]
    taps = _fir_axis_taps(f.to(x.dtype) * (gain ** 0.5), flip)
    k = taps.numel()
    x = F.conv2d(x, taps.reshape(1, 1, 1, k).repeat(c, 1, 1, 1), groups=c)      # along x
    x = F.conv2d(x, taps.reshape(1, 1, k, 1).repeat(c, 1, 1, 1), groups=c)      # along y
    return x[:, :, ::down, ::down]


def fir_upsample2(x, f):
    """upfirdn2d.upsample2d(x, f, up=2) (`ops/upfirdn2d.py:300-337`): gain up^2, padding such that
    the output is exactly twice the input."""
    k = f.numel()
    p = ((k + 1) // 2, (k - 2) // 2)
    return fir_resample(x, f, up=2, pad=(p[0], p[1], p[0], p[1]), gain=4.0)


def fir_downsample2(x, f, padding=0, flip=False):
    """upfirdn2d.downsample2d(x, f, down=2, padding, flip_filter) (`ops/upfirdn2d.py:341-378`)."""
    k = f.numel()
    p = (padding + (k - 1) // 2, padding + (k - 2) // 2)
    return fir_resample(x, f, down=2, pad=(p[0], p[1], p[0], p[1]), flip=flip)


# ------------------------------------------------------------------ sampled parameters
@dataclass
class AugmentParams:
    """One draw of every per-image random quantity of the pipe (None = stage not active)."""
    G_inv: Optional[torch.Tensor] = None        # [B,3,3]  output pixel -> input pixel (centred coords)
    C: Optional[torch.Tensor] = None            # [B,4,4]  homogeneous colour transform
    band_gain: Optional[torch.Tensor] = None    # [B,4]    gains of the image-space filter bands
    noise_sigma: Optional[torch.Tensor] = None  # [B,1,1,1]
    noise_field: Optional[torch.Tensor] = None  # [B,C,H,W] unit normal
    cut_size: Optional[torch.Tensor] = None     # [B,2,1,1,1]
    cut_center: Optional[torch.Tensor] = None   # [B,2,1,1,1]
    shape: tuple = field(default_factory=tuple)


def _gate(rand_like_shape, prob, value, neutral, device):
    """`value` where a uniform draw falls below `prob`, else `neutral` (one torch.rand call)."""
    return torch.where(torch.rand(rand_like_shape, device=device) < prob, value, neutral)


class AugmentPipe(torch.nn.Module):
    """All augmentations are off by default; a stage is switched on by giving it a probability
    multiplier (ada/augment.py:118-180, same arguments).  `p` (buffer) is the overall strength."""

    def __init__(self, xflip=0, rotate90=0, xint=0, xint_max=0.125,
                 scale=0, rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2,
                 xfrac_std=0.125,
                 brightness=0, contrast=0, lumaflip=0, hue=0, saturation=0, brightness_std=0.2,
                 contrast_std=0.5, hue_max=1, saturation_std=1,
                 imgfilter=0, imgfilter_bands=(1, 1, 1, 1), imgfilter_std=1,
                 noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        self.register_buffer("p", torch.ones([]))
        for name, val in dict(xflip=xflip, rotate90=rotate90, xint=xint, xint_max=xint_max, scale=scale,
                              rotate=rotate, aniso=aniso, xfrac=xfrac, scale_std=scale_std, rotate_max=rotate_max,
                              aniso_std=aniso_std, xfrac_std=xfrac_std, brightness=brightness, contrast=contrast,
                              lumaflip=lumaflip, hue=hue, saturation=saturation, brightness_std=brightness_std,
                              contrast_std=contrast_std, hue_max=hue_max, saturation_std=saturation_std,
                              imgfilter=imgfilter, imgfilter_std=imgfilter_std, noise=noise, cutout=cutout,
                              noise_std=noise_std, cutout_size=cutout_size).items():
            setattr(self, name, float(val))
        self.imgfilter_bands = list(imgfilter_bands)
        lo = torch.tensor(_SYM6, dtype=torch.float32)
        self.register_buffer("Hz_geom", lo / lo.sum())                  # setup_filter(): DC gain 1
        self.register_buffer("Hz_fbank", torch.as_tensor(self._band_filters(), dtype=torch.float32))

    @staticmethod
    def _band_filters():
        """4 x 13 bank of zero-phase band filters from the sym2 pair (ada/augment.py:167-177):
        row 0 = lowest band ... row 3 = highest; their sum is the unit impulse."""
        lo = np.asarray(_SYM2)
        hi = lo * ((-1) ** np.arange(lo.size))
        lo2 = np.convolve(lo, lo[::-1]) / 2
        hi2 = np.convolve(hi, hi[::-1]) / 2
        bank = np.eye(4, 1)
        for i in range(1, 4):
            bank = np.dstack([bank, np.zeros_like(bank)]).reshape(4, -1)[:, :-1]      # zero-stuff x2
            bank = np.stack([np.convolve(row, lo2) for row in bank])
            mid = bank.shape[1] // 2
            bank[i, mid - hi2.size // 2: mid - hi2.size // 2 + hi2.size] += hi2
        return bank

    # -------------------------------------------------------------- sampling
    def sample(self, batch_size, num_channels, height, width, device, debug_percentile=None):
        """Draw every random quantity, in the reference's order (ada/augment.py:182-428)."""
        B, dev, pct = batch_size, device, debug_percentile
        if pct is not None:
            pct = torch.as_tensor(pct, dtype=torch.float32, device=dev)
        P = AugmentParams(shape=(B, num_channels, height, width))
        ones = torch.ones([B], device=dev)
        zeros = torch.zeros([B], device=dev)

        def icdf(std):          # debug: the `pct` quantile of N(0, std)
            return torch.erfinv(pct * 2 - 1) * std

        # ---- geometry: G_inv @ (output pixel) = input pixel
        G = None

        def then(M):
            nonlocal G
            G = M if G is None else G @ M

        if self.xflip > 0:
            i = torch.floor(torch.rand([B], device=dev) * 2)
            i = _gate([B], self.xflip * self.p, i, zeros, dev)
            if pct is not None:
                i = torch.full_like(i, torch.floor(pct * 2))
            then(_scale2(1 / (1 - 2 * i), ones))
        if self.rotate90 > 0:
            i = torch.floor(torch.rand([B], device=dev) * 4)
            i = _gate([B], self.rotate90 * self.p, i, zeros, dev)
            if pct is not None:
                i = torch.full_like(i, torch.floor(pct * 4))
            then(_rot2(math.pi / 2 * i))
        if self.xint > 0:
            t = (torch.rand([B, 2], device=dev) * 2 - 1) * self.xint_max
            t = _gate([B, 1], self.xint * self.p, t, torch.zeros_like(t), dev)
            if pct is not None:
                t = torch.full_like(t, (pct * 2 - 1) * self.xint_max)
            then(_shift2(-torch.round(t[:, 0] * width), -torch.round(t[:, 1] * height)))
        if self.scale > 0:
            s = torch.exp2(torch.randn([B], device=dev) * self.scale_std)
            s = _gate([B], self.scale * self.p, s, ones, dev)
            if pct is not None:
                s = torch.full_like(s, torch.exp2(icdf(self.scale_std)))
            then(_scale2(1 / s, 1 / s))
        p_rot = 1 - torch.sqrt((1 - self.rotate * self.p).clamp(0, 1))      # P(pre or post) = rotate * p
        if self.rotate > 0:
            th = (torch.rand([B], device=dev) * 2 - 1) * math.pi * self.rotate_max
            th = _gate([B], p_rot, th, zeros, dev)
            if pct is not None:
                th = torch.full_like(th, (pct * 2 - 1) * math.pi * self.rotate_max)
            then(_rot2(th))
        if self.aniso > 0:
            s = torch.exp2(torch.randn([B], device=dev) * self.aniso_std)
            s = _gate([B], self.aniso * self.p, s, ones, dev)
            if pct is not None:
                s = torch.full_like(s, torch.exp2(icdf(self.aniso_std)))
            then(_scale2(1 / s, 1 / (1 / s)))
        if self.rotate > 0:
            th = (torch.rand([B], device=dev) * 2 - 1) * math.pi * self.rotate_max
            th = _gate([B], p_rot, th, zeros, dev)
            if pct is not None:
                th = torch.zeros_like(th)
            then(_rot2(th))
        if self.xfrac > 0:
            t = torch.randn([B, 2], device=dev) * self.xfrac_std
            t = _gate([B, 1], self.xfrac * self.p, t, torch.zeros_like(t), dev)
            if pct is not None:
                t = torch.full_like(t, icdf(self.xfrac_std))
            then(_shift2(-(t[:, 0] * width), -(t[:, 1] * height)))
        P.G_inv = G

        # ---- colour: C @ (r, g, b, 1)
        C = None

        def before(M):
            nonlocal C
            C = M if C is None else M @ C

        if self.brightness > 0:
            b = torch.randn([B], device=dev) * self.brightness_std
            b = _gate([B], self.brightness * self.p, b, zeros, dev)
            if pct is not None:
                b = torch.full_like(b, icdf(self.brightness_std))
            before(_shift3(b, b, b))
        if self.contrast > 0:
            c = torch.exp2(torch.randn([B], device=dev) * self.contrast_std)
            c = _gate([B], self.contrast * self.p, c, ones, dev)
            if pct is not None:
                c = torch.full_like(c, torch.exp2(icdf(self.contrast_std)))
            before(_scale3(c, c, c))
        luma = torch.tensor(np.asarray([1, 1, 1, 0]) / np.sqrt(3), dtype=torch.float32, device=dev)      # grey axis
        eye4 = torch.eye(4, device=dev)
        if self.lumaflip > 0:
            i = torch.floor(torch.rand([B, 1, 1], device=dev) * 2)
            i = _gate([B, 1, 1], self.lumaflip * self.p, i, torch.zeros_like(i), dev)
            if pct is not None:
                i = torch.full_like(i, torch.floor(pct * 2))
            before(eye4 - 2 * torch.outer(luma, luma) * i)                 # Householder about the grey axis
        if self.hue > 0 and num_channels > 1:
            th = (torch.rand([B], device=dev) * 2 - 1) * math.pi * self.hue_max
            th = _gate([B], self.hue * self.p, th, zeros, dev)
            if pct is not None:
                th = torch.full_like(th, (pct * 2 - 1) * math.pi * self.hue_max)
            before(_rot3(luma, th))
        if self.saturation > 0 and num_channels > 1:
            s = torch.exp2(torch.randn([B, 1, 1], device=dev) * self.saturation_std)
            s = _gate([B, 1, 1], self.saturation * self.p, s, torch.ones_like(s), dev)
            if pct is not None:
                s = torch.full_like(s, torch.exp2(icdf(self.saturation_std)))
            ll = torch.outer(luma, luma)
            before(ll + (eye4 - ll) * s)
        P.C = C

        # ---- image-space filter: per-band gains, power-normalised one band at a time
        if self.imgfilter > 0:
            power = torch.tensor(np.array([10, 1, 1, 1]) / 13, dtype=torch.float32, device=dev)      # expected 1/f spectrum
            g = torch.ones([B, 4], device=dev)
            for i, strength in enumerate(self.imgfilter_bands):
                t_i = torch.exp2(torch.randn([B], device=dev) * self.imgfilter_std)
                t_i = _gate([B], self.imgfilter * self.p * strength, t_i, ones, dev)
                if pct is not None:
                    t_i = torch.full_like(t_i, torch.exp2(icdf(self.imgfilter_std))) if strength > 0 \
                        else torch.ones_like(t_i)
                t = torch.ones([B, 4], device=dev)
                t[:, i] = t_i
                t = t / (power * t.square()).sum(dim=-1, keepdims=True).sqrt()
                g = g * t
            P.band_gain = g

        # ---- corruptions
        if self.noise > 0:
            sigma = torch.randn([B, 1, 1, 1], device=dev).abs() * self.noise_std
            sigma = _gate([B, 1, 1, 1], self.noise * self.p, sigma, torch.zeros_like(sigma), dev)
            if pct is not None:
                sigma = torch.full_like(sigma, torch.erfinv(pct) * self.noise_std)
            P.noise_sigma = sigma
            P.noise_field = torch.randn([B, num_channels, height, width], device=dev)
        if self.cutout > 0:
            size = torch.full([B, 2, 1, 1, 1], self.cutout_size, device=dev)
            size = _gate([B, 1, 1, 1, 1], self.cutout * self.p, size, torch.zeros_like(size), dev)
            center = torch.rand([B, 2, 1, 1, 1], device=dev)
            if pct is not None:
                size = torch.full_like(size, self.cutout_size)
                center = torch.full_like(center, pct)
            P.cut_size, P.cut_center = size, center
        return P

    # -------------------------------------------------------------- execution
    def _warp(self, images, G_inv):
        """Reflect-pad by the margin the transform can reach, x2 sym6 upsampling, bilinear
        resampling by the affine map, sym6 x2 downsampling + crop (ada/augment.py:262-295)."""
        B, Cc, H, W = images.shape
        dev = images.device
        f = self.Hz_geom
        cx, cy = (W - 1) / 2, (H - 1) / 2
        corners = torch.tensor([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], dtype=torch.float32, device=dev)
        reach = G_inv @ corners.t()                                       # [B, xyz, corner]
        halo = f.shape[0] // 4
        m = reach[:, :2, :].permute(1, 0, 2).flatten(1)                   # [xy, B * corner]
        m = torch.cat([-m, m]).max(dim=1).values                          # x0, y0, x1, y1
        m = m + torch.tensor([halo * 2 - cx, halo * 2 - cy] * 2, dtype=torch.float32, device=dev)
        m = m.max(torch.zeros(4, device=dev))
        m = m.min(torch.tensor([W - 1, H - 1] * 2, dtype=torch.float32, device=dev))
        mx0, my0, mx1, my1 = (int(v) for v in m.ceil().to(torch.int32))
        x = F.pad(images, [mx0, mx1, my0, my1], mode="reflect")
        G = _shift2((mx0 - mx1) / 2, (my0 - my1) / 2, device=dev) @ G_inv
        x = fir_upsample2(x, f)
        G = _scale2(2, 2, device=dev) @ G @ _scale2(1 / 2, 1 / 2, device=dev)
        G = _shift2(-0.5, -0.5, device=dev) @ G @ _shift2(0.5, 0.5, device=dev)
        out_shape = [B, Cc, (H + halo * 2) * 2, (W + halo * 2) * 2]
        G = _scale2(2 / x.shape[3], 2 / x.shape[2], device=dev) @ G @ \
            _scale2(1 / (2 / out_shape[3]), 1 / (2 / out_shape[2]), device=dev)
        grid = F.affine_grid(theta=G[:, :2, :], size=out_shape, align_corners=False)
        x = _Resample.apply(x, grid.to(x.dtype))
        return fir_downsample2(x, f, padding=-halo * 2, flip=True)

    def apply(self, images, P):
        """Run the sampled transforms on `images` [B,C,H,W] (differentiable any number of times
        w.r.t. `images`)."""
        B, Cc, H, W = images.shape
        if P.G_inv is not None:
            images = self._warp(images, P.G_inv.float())        # geometry in fp32, as the reference
        if P.C is not None:
            C = P.C.to(images.dtype)
            flat = images.reshape(B, Cc, H * W)
            if Cc == 3:
                flat = C[:, :3, :3] @ flat + C[:, :3, 3:]
            elif Cc == 1:
                C1 = C[:, :3, :].mean(dim=1, keepdims=True)
                flat = flat * C1[:, :, :3].sum(dim=2, keepdims=True) + C1[:, :, 3:]
            else:
                raise ValueError("Image must be RGB (3 channels) or L (1 channel)")
            images = flat.reshape(B, Cc, H, W)
        if P.band_gain is not None:
            taps = (P.band_gain.to(images.dtype) @ self.Hz_fbank.to(images.dtype))          # [B, 13]
            taps = taps.unsqueeze(1).repeat(1, Cc, 1).reshape(B * Cc, 1, -1)
            half = self.Hz_fbank.shape[1] // 2
            x = images.reshape(1, B * Cc, H, W)
            x = F.pad(x, [half, half, half, half], mode="reflect")
            x = F.conv2d(x, taps.unsqueeze(2), groups=B * Cc)
            x = F.conv2d(x, taps.unsqueeze(3), groups=B * Cc)
            images = x.reshape(B, Cc, H, W)
        if P.noise_sigma is not None:
            images = images + P.noise_field.to(images.dtype) * P.noise_sigma.to(images.dtype)
        if P.cut_size is not None:
            dev = images.device
            u = (torch.arange(W, device=dev).reshape(1, 1, 1, -1) + 0.5) / W
            v = (torch.arange(H, device=dev).reshape(1, 1, -1, 1) + 0.5) / H
            keep_x = (u - P.cut_center[:, 0]).abs() >= P.cut_size[:, 0] / 2
            keep_y = (v - P.cut_center[:, 1]).abs() >= P.cut_size[:, 1] / 2
            images = images * torch.logical_or(keep_x, keep_y).to(images.dtype)
        return images

    def forward(self, images, debug_percentile=None):
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        B, Cc, H, W = images.shape
        return self.apply(images, self.sample(B, Cc, H, W, images.device, debug_percentile))


class AdaptiveAugment:
    """ADA-p controller (ada/adapt_augm.py:6-51): p moves by `batch / ada_length` per evaluated
    image towards keeping r_t = E[sign(D(real))] at `ada_target`, re-evaluated every 4 batches.
    Same arithmetic and return values; the running sign sum stays on the device and is read on
    the host only when a window closes (the reference syncs with .item() on every update)."""

    def __init__(self, prev_ada_p=0.0, ada_target=0.6, ada_length=500000, batch_size=4, device="cpu"):
        self.prev_ada_p, self.ada_target, self.ada_length = prev_ada_p, ada_target, ada_length
        self.batch_size, self.rank = batch_size, device
        self.ada_aug_step = 1.0 / (self.ada_length / self.batch_size)

    def initialize(self):
        self._sign_sum = torch.zeros((), device=self.rank)
        self._count = 0
        self.ada_aug_p = self.prev_ada_p if self.prev_ada_p is not None else 0.0
        return self.ada_aug_p

    def update(self, logits):
        self._sign_sum = self._sign_sum + torch.sign(logits).sum().to(self._sign_sum.dtype)
        self._count += logits.shape[0]
        if self._count > self.batch_size * 4 - 1:
            r_t = float(self._sign_sum) / self._count
            direction = 1 if r_t > self.ada_target else -1
            self.ada_aug_p = min(1.0, max(0.0, self.ada_aug_p + direction * self.ada_aug_step * self._count))
            self._sign_sum = torch.zeros_like(self._sign_sum)
            self._count = 0
        return self.ada_aug_p

    def set_batch_size(self, batch_size):
        self.batch_size = batch_size
        self.ada_aug_step = 1.0 / (self.ada_length / self.batch_size)
