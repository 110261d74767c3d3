"""TEST DOUBLE for the CUDA backend (tests only — never imported by the product).

Implements the *semantics* of every C-ABI kernel (include/progan_b200.h) with torch ops on
any device/dtype, behind the same method names as progan_b200.kernels.CudaKernels.  It lets
the `-m "not gpu"` suite exercise the host logic — module wiring, the autograd Function
families, the hand-derived second-order formulas — on a machine without a GPU, and serves
as the per-kernel specification the GPU unit tests compare against.
"""
import torch
import torch.nn.functional as F

from progan_b200.kernels import EPI_LINEAR, EPI_LRELU, EPI_PN_LRELU


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def logical_weight(w, op):
    wl = w.transpose(0, 1) if op.swap else w
    if op.flip:
        wl = wl.flip(2, 3)
    return wl


class EmulKernels:
    name = "emul"

    def __init__(self):
        self.launches = 0
        self.conv_impl = "simt"

    defer_wgrad = False
    wgrad_side_stream = None

    def invalidate_packs(self):
        pass

    def refresh_packs(self, params):
        pass

    def drop_packs(self, params):
        pass

    def flush_wgrads(self):
        pass

    def mbstd_channels(self, C, dtype):
        return C + 1

    # ---- conv
    def conv_fwd(self, x, w, bias, op, scale, epi=EPI_LINEAR, slope=0.2, pool_out=False):
        self.launches += 1
        cd = torch.float64 if x.dtype == torch.float64 else torch.float32
        wl = logical_weight(w, op).to(x.dtype).to(cd)   # packed operands carry the activation dtype
        x = x[..., :op.cin(w.shape)]                    # padded channels carry zero weights
        a = F.conv2d(_nchw(x).to(cd), wl, None, padding=op.pad) * scale
        if bias is not None:
            a = a + bias.to(cd).view(1, -1, 1, 1)
        r = None
        if epi == EPI_PN_LRELU:
            r = torch.rsqrt((a * a).mean(dim=1) + 1e-8)
            a = a * r.unsqueeze(1)
        if epi != EPI_LINEAR:
            a = torch.where(a > 0, a, a * slope)
        y = _nhwc(a).to(x.dtype)
        if op.ypad and op.ypad > y.shape[-1]:
            y = F.pad(y, (0, op.ypad - y.shape[-1]))
        if r is not None:
            r = r.contiguous().to(torch.float64 if cd == torch.float64 else torch.float32)
        if pool_out:
            return y, r, None          # the emulation pools with the separate kernel
        return y, r

    def conv_wgrad(self, x, dy, wshape, op, scale, out=None):
        self.launches += 1
        cd = torch.float64 if x.dtype == torch.float64 else torch.float32
        cout, cin = op.cout(wshape), op.cin(wshape)
        x, dy = x[..., :cin], dy[..., :cout]
        dwl = torch.nn.grad.conv2d_weight(_nchw(x).to(cd), (cout, cin, op.k, op.k),
                                          _nchw(dy).to(cd), padding=op.pad) * scale
        if op.flip:
            dwl = dwl.flip(2, 3)
        if op.swap:
            dwl = dwl.transpose(0, 1)
        res = dwl.contiguous().to(cd if cd == torch.float64 else torch.float32)
        if out is not None:
            out.add_(res.to(out.dtype))
            return out
        return res

    # ---- PixelNorm + LeakyReLU
    @staticmethod
    def _pm(y, slope):
        pos = y > 0
        m = torch.where(pos, torch.ones_like(y), torch.full_like(y, slope))
        p = torch.where(pos, y, y / slope)
        return p, m

    @staticmethod
    def _unpool(dy):
        return 0.25 * dy.repeat_interleave(2, 1).repeat_interleave(2, 2)

    def pn_lrelu_bwd(self, dy, y, r, slope, use_pn, pool=False, want_colsum=False, colsum_out=None,
                     addend=None):
        self.launches += 1
        cd = torch.float64 if y.dtype == torch.float64 else torch.float32
        p, m = self._pm(y.to(cd), slope)
        dyf = self._unpool(dy.to(cd)) if pool else dy.to(cd)
        u = m * dyf
        if use_pn:
            C = y.shape[-1]
            s = (p * u).sum(-1, keepdim=True)
            u = r.to(cd).unsqueeze(-1) * (u - p * s / C)
        if addend is not None:
            u = u + addend.to(cd)
        cs = None
        if want_colsum or colsum_out is not None:
            cs = u.reshape(-1, y.shape[-1]).sum(0).to(cd if cd == torch.float64 else torch.float32)
            if colsum_out is not None:
                colsum_out.add_(cs.to(colsum_out.dtype))
                cs = colsum_out
        return u.to(y.dtype), cs

    def pn_lrelu_bwd_bwd(self, t, dy, y, r, slope, use_pn, pool=False):
        self.launches += 1
        cd = torch.float64 if y.dtype == torch.float64 else torch.float32
        p, m = self._pm(y.to(cd), slope)
        dyf = self._unpool(dy.to(cd)) if pool else dy.to(cd)
        u = m * dyf
        t = t.to(cd)
        if not use_pn:
            return (m * t).to(y.dtype), torch.zeros_like(y)
        C = y.shape[-1]
        rr = r.to(cd).unsqueeze(-1)
        s_pt = (p * t).sum(-1, keepdim=True)
        s_pu = (p * u).sum(-1, keepdim=True)
        s_tu = (t * u).sum(-1, keepdim=True)
        cot_dy = m * rr * (t - p * s_pt / C)
        cot_a = rr * rr / C * (3.0 / C * s_pt * s_pu * p - s_tu * p - s_pu * t - s_pt * u)
        return cot_dy.to(y.dtype), cot_a.to(y.dtype)

    def colsum(self, x, out=None):
        self.launches += 1
        res = x.reshape(-1, x.shape[-1]).to(torch.float64 if x.dtype == torch.float64 else torch.float32).sum(0)
        if out is not None:
            out.add_(res.to(out.dtype))
            return out
        return res

    # ---- 1x1 heads
    @staticmethod
    def _wck(w, C, Kc, w_sc, w_sk):
        flat = w.reshape(-1)
        idx = (torch.arange(C, device=w.device).view(C, 1) * w_sc
               + torch.arange(Kc, device=w.device).view(1, Kc) * w_sk)
        return flat[idx]      # [C, K]

    def pw_expand(self, img, w, bias, C, w_sc, w_sk, scale, dtype):
        self.launches += 1
        Kc = img.shape[1]
        wck = self._wck(w, C, Kc, w_sc, w_sk).to(img.dtype)
        out = torch.einsum('nkhw,ck->nhwc', img, wck) * scale
        if bias is not None:
            out = out + bias.to(img.dtype)
        return out.contiguous().to(dtype)

    def pw_reduce(self, act, w, bias, Kc, w_sc, w_sk, scale):
        self.launches += 1
        C = act.shape[-1]
        cd = torch.float64 if act.dtype == torch.float64 else torch.float32
        wck = self._wck(w, C, Kc, w_sc, w_sk).to(cd)
        out = torch.einsum('nhwc,ck->nkhw', act.to(cd), wck) * scale
        if bias is not None:
            out = out + bias.to(cd).view(1, -1, 1, 1)
        return out.contiguous()

    def pw_wgrad(self, act, img, wshape, w_sc, w_sk, scale, out=None, bias_out=None):
        self.launches += 1
        C, Kc = act.shape[-1], img.shape[1]
        cd = img.dtype
        if bias_out is not None:
            bias_out.add_(act.to(cd).reshape(-1, C).sum(0).to(bias_out.dtype))
        dck = torch.einsum('nhwc,nkhw->ck', act.to(cd), img) * scale
        dw = torch.zeros(wshape, dtype=cd, device=img.device)
        idx = (torch.arange(C, device=img.device).view(C, 1) * w_sc
               + torch.arange(Kc, device=img.device).view(1, Kc) * w_sk)
        dw.view(-1)[idx.reshape(-1)] = dck.reshape(-1)
        if out is not None:
            out.add_(dw.to(out.dtype))
            return out
        return dw

    def img_chansum(self, img, out=None):
        self.launches += 1
        res = img.sum(dim=(0, 2, 3))
        if out is not None:
            out.add_(res.to(out.dtype))
            return out
        return res

    # ---- resampling
    @staticmethod
    def _as_nchw(x, fmt):
        return x if fmt == "nchw" else _nchw(x)

    @staticmethod
    def _back(y, fmt, dtype):
        return (y if fmt == "nchw" else _nhwc(y)).contiguous().to(dtype)

    def _cd(self, x):
        return torch.float64 if x.dtype == torch.float64 else torch.float32

    def avgpool2(self, x, fmt="nhwc"):
        self.launches += 1
        return self._back(F.avg_pool2d(self._as_nchw(x, fmt).to(self._cd(x)), 2), fmt, x.dtype)

    def avgpool2_bwd(self, dy, fmt="nhwc"):
        self.launches += 1
        g = self._as_nchw(dy, fmt).to(self._cd(dy))
        return self._back(0.25 * g.repeat_interleave(2, 2).repeat_interleave(2, 3), fmt, dy.dtype)

    def upsample2(self, x, fmt="nhwc"):
        self.launches += 1
        y = F.interpolate(self._as_nchw(x, fmt).to(self._cd(x)), scale_factor=2, mode='bilinear',
                          align_corners=False)
        return self._back(y, fmt, x.dtype)

    def upsample2_bwd(self, dy, fmt="nhwc"):
        self.launches += 1
        g = self._as_nchw(dy, fmt).to(self._cd(dy))
        N, C, H2, W2 = g.shape
        x = torch.zeros(N, C, H2 // 2, W2 // 2, dtype=g.dtype, device=g.device, requires_grad=True)
        with torch.enable_grad():
            y = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
            (dx,) = torch.autograd.grad(y, x, g)
        return self._back(dx, fmt, dy.dtype)

    # ---- elementwise
    def blend(self, a, b, alpha_dev):
        self.launches += 1
        al = alpha_dev.to(self._cd(a))
        return ((1 - al) * a.to(self._cd(a)) + al * b.to(self._cd(a))).to(a.dtype)

    def scale(self, x, c0, c1, alpha_dev):
        self.launches += 1
        coef = c0 + (c1 * alpha_dev.to(self._cd(x)) if alpha_dev is not None else 0.0)
        return (coef * x.to(self._cd(x))).to(x.dtype)

    def tanh_fwd(self, x):
        self.launches += 1
        return torch.tanh(x)

    def tanh_bwd(self, dy, y):
        self.launches += 1
        return dy * (1 - y * y)

    # ---- mbstd   x:[N,4,4,C]
    def _mb_stats(self, x):
        xf = x.to(self._cd(x))
        mu = xf.mean(0, keepdim=True)
        sigma = torch.sqrt(((xf - mu) ** 2).mean(0, keepdim=True) + 1e-8)
        return xf, mu, sigma

    def mbstd_fwd(self, x, Cp):
        self.launches += 1
        xf, mu, sigma = self._mb_stats(x)
        N, _, _, C = x.shape
        out = torch.zeros(N, 4, 4, Cp, dtype=xf.dtype, device=x.device)
        out[..., :C] = xf
        out[..., C] = sigma.mean()
        return out.to(x.dtype), None

    def mbstd_bwd(self, dout, x, stats=None):
        self.launches += 1
        xf, mu, sigma = self._mb_stats(x)
        N, _, _, C = x.shape
        Fn = 16 * C
        g = dout.to(xf.dtype)
        dm = g[..., C].sum()
        return (g[..., :C] + dm * (xf - mu) / (N * Fn * sigma)).contiguous().to(x.dtype)

    def mbstd_bwd_bwd(self, t, dout, x, stats=None):
        self.launches += 1
        xf, mu, sigma = self._mb_stats(x)
        N, _, _, C = x.shape
        Fn = 16 * C
        g = dout.to(xf.dtype)
        tt = t.to(xf.dtype)
        dm = g[..., C].sum()
        xc = xf - mu
        tbar = tt.mean(0, keepdim=True)
        cf = (tt * xc).mean(0, keepdim=True)
        tau = (cf / (Fn * sigma)).sum()
        cot_x = dm / (N * Fn * sigma) * (tt - tbar - xc * cf / sigma ** 2)
        cot_dout = torch.zeros_like(g)
        cot_dout[..., :C] = tt
        cot_dout[..., C] = tau
        return cot_dout.to(dout.dtype), cot_x.contiguous().to(x.dtype)

    # ---- WGAN-GP
    def wgan_loss(self, d, n_real, drift, metric=None):
        self.launches += 1
        flat = d.reshape(-1)
        n = flat.numel()
        seed = torch.empty_like(flat)
        if n_real > 0:
            dr, df = flat[:n_real], flat[n_real:]
            seed[:n_real] = (-1.0 + 2.0 * drift * dr) / n_real
            seed[n_real:] = 1.0 / (n - n_real)
            val = dr.mean() - drift * (dr * dr).mean() - df.mean()
        else:
            seed[:] = -1.0 / n
            val = -flat.mean()
        if metric is not None:
            metric.add_(val.to(metric.dtype))
        return seed.view_as(d)

    def interp_xhat(self, real, fake, eps):
        self.launches += 1
        e = eps.view(-1, *([1] * (real.dim() - 1)))
        return e * real + (1 - e) * fake

    def gp_fwd(self, g, lam):
        self.launches += 2
        norms = g.reshape(g.shape[0], -1).norm(2, dim=1)
        return lam * ((norms - 1) ** 2).mean(), norms

    def gp_bwd(self, g, norms, upstream, lam):
        self.launches += 1
        N = g.shape[0]
        k = upstream * 2 * lam / N * (norms - 1) / norms
        return g * k.view(-1, *([1] * (g.dim() - 1)))

    # ---- optimiser
    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, step_dev, grad_scale=1.0):
        self.launches += 1
        t = float(step_dev)
        gi = g * grad_scale
        mi = gi
        if m is not None:
            m.mul_(beta1).add_(gi, alpha=1 - beta1)
            mi = m
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        denom = v.sqrt() / (1 - beta2 ** t) ** 0.5 + eps
        p.sub_(lr / (1 - beta1 ** t) * mi / denom)

    def adam_multi(self, p, g, m, v, chunks, steps_dev, lr, beta1, beta2, eps, grad_scale=1.0):
        self.launches += 1
        for start, length, seg, _ in chunks.tolist():
            sl = slice(start, start + length)
            self.adam_step(p[sl], g[sl], None if m is None else m[sl], v[sl], lr, beta1, beta2, eps,
                           steps_dev[seg], grad_scale)
            self.launches -= 1

    def ema(self, ema, p, decay):
        self.launches += 1
        ema.mul_(decay).add_(p, alpha=1 - decay)
