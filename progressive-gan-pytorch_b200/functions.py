"""torch.autograd.Function families over the explicit kernels.

The reference gets every derivative of the step from autograd over ATen ops, including
the WGAN-GP double backward (train.py:146-151).  Here each fused op is a Function whose
backward is *another* Function — the pattern the reference's own plugin ops use
(ada/torch_utils/ops/bias_act.py:145-206, conv2d_gradfix.py:107-165) — so that
`torch.autograd.grad(..., create_graph=True)` followed by `.backward()` in the train
scripts keeps working, but every node of both graphs is one explicit kernel:

  conv family   : ConvFwd(op) <-> ConvFwd(op.adjoint()) (data-grad), ConvWgrad(op)
  1x1 heads     : PwFwd(expand|reduce), PwWgrad
  activation    : ConvAct (fused conv+bias+PixelNorm+LeakyReLU epilogue) -> Act -> ActBwd,
                  ActBwd.backward = the hand-derived PixelNorm second-order kernel
  resampling    : Linear1(avgpool2|avgpool2_bwd|upsample2|upsample2_bwd), Blend, Scale
  mbstd, tanh, gradient-penalty scalar.

The tensor that links ConvAct to Act holds the stored post-activation y, but autograd-wise
it *is* the pre-activation a: gradients w.r.t. it are true d/da, so the Hessian term that
PixelNorm contributes during the GP double backward and the ordinary chain-rule term are
summed by autograd and sent through ONE data-grad + ONE weight-grad per conv (6 F_D total,
SURVEY.md §3.4), and the pre-activation never has to be written to HBM.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as _k
from .kernels import ConvOp, EPI_LINEAR, EPI_LRELU, EPI_PN_LRELU


def K():
    return _k.get_kernels()


def _acc_node(p):
    """The AccumulateGrad node of a leaf parameter.  Deliberately NOT cached: a node kept alive
    across iterations keeps the stream it was created on, which breaks CUDA-graph capture
    (the engine would sync the capturing stream with that stale stream).  Grad mode is forced
    on: inside a plain backward it is off and view_as() would not record a node."""
    with torch.enable_grad():
        return p.view_as(p).grad_fn.next_functions[0][0]


# Direct gradient accumulation (enabled by Trainer): weight/bias gradient kernels add straight
# into the parameter's .grad (a view of the flat bucket) and the Function returns None, which
# removes autograd's AccumulateGrad add kernels and the zero-initialised temporaries.  Only
# for plain (non-create_graph) backward passes of leaf parameters that already own a .grad.
DIRECT_GRADS = False


def _direct_target(p):
    if (DIRECT_GRADS and p is not None and not torch.is_grad_enabled() and p.is_leaf
            and p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32):
        return p.grad
    return None


class ActLink:
    """Ties an Act node to the ONE conv that consumes its output (the second conv of a
    ConvBlock).  In a plain (non-create_graph) backward that conv's data-gradient kernel applies
    this activation's backward in its epilogue and sets `fused`; the Act node then passes the
    incoming gradient through unchanged."""
    __slots__ = ("A", "r", "slope", "use_pn", "bias", "fused")

    def __init__(self, A, r, slope, use_pn, bias):
        self.A, self.r, self.slope, self.use_pn, self.bias = A, r, slope, use_pn, bias
        self.fused = False


FUSE_ACT_BWD = True


def _wants_grad(ctx, idx, t):
    """needs_input_grad refined by the engine's execution plan: for
    autograd.grad(inputs=[x_hat]) / backward(inputs=G.parameters()) the weight-gradient
    kernels of parameters that are not requested are skipped, as ATen's own
    convolution_backward does through its output mask."""
    if not ctx.needs_input_grad[idx]:
        return False
    if t is None or not (t.is_leaf and t.requires_grad):
        return True
    try:
        return bool(torch._C._will_engine_execute_node(_acc_node(t)))
    except Exception:
        return True


# --------------------------------------------------------------------------- conv
class ConvFwd(Function):
    """y = scale * conv(x; Wl(w)) — linear in x and w, no bias."""

    @staticmethod
    def forward(ctx, x, w, op, scale):
        ctx.op, ctx.scale = op, scale
        ctx.save_for_backward(x, w)
        y, _ = K().conv_fwd(x, w, None, op, scale, EPI_LINEAR)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ConvFwd.apply(dy, w, ctx.op.adjoint(), ctx.scale)
        if _wants_grad(ctx, 1, w):
            tgt = _direct_target(w)
            if tgt is not None:
                K().conv_wgrad(x, dy, tuple(w.shape), ctx.op, ctx.scale, out=tgt)
            else:
                dw = ConvWgrad.apply(x, dy, ctx.op, ctx.scale, tuple(w.shape))
        return dx, dw, None, None


class ConvWgrad(Function):
    """dw = scale * sum_pix dy (x) x  in the parameter's own layout."""

    @staticmethod
    def forward(ctx, x, dy, op, scale, wshape):
        ctx.op, ctx.scale = op, scale
        ctx.save_for_backward(x, dy)
        return K().conv_wgrad(x, dy, wshape, op, scale)

    @staticmethod
    def backward(ctx, ddw):
        x, dy = ctx.saved_tensors
        ddw = ddw.contiguous()
        cx = cdy = None
        if ctx.needs_input_grad[0]:
            cx = ConvFwd.apply(dy, ddw, ctx.op.adjoint(), ctx.scale)
        if ctx.needs_input_grad[1]:
            cdy = ConvFwd.apply(x, ddw, ctx.op, ctx.scale)
        return cx, cdy, None, None, None


class ColSum(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return K().colsum(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return g.to(ctx.dtype).expand(ctx.shape).contiguous()


class ConvAct(Function):
    """Fused y = lrelu(pixelnorm(scale*conv(x;w) + b)) — one kernel, epilogue-fused.
    Returns (A, r): A carries y's data but stands for the pre-activation in the graph."""

    @staticmethod
    def forward(ctx, x, w, b, op, scale, slope, use_pn, prev_link=None, pool=False):
        ctx.op, ctx.scale = op, scale
        ctx.prev_link = prev_link
        ctx.save_for_backward(x, w, b)
        ctx.set_materialize_grads(False)     # r never gets a gradient: no zero-fill per backward
        epi = EPI_PN_LRELU if use_pn else EPI_LRELU
        yp = None
        if pool:          # the kernel also writes avgpool2(y) where its epilogue can (else None)
            y, r, yp = K().conv_fwd(x, w, b, op, scale, epi, slope, pool_out=True)
        else:
            y, r = K().conv_fwd(x, w, b, op, scale, epi, slope)
        if r is None:
            r = torch.empty(0, device=x.device, dtype=torch.float32)
        if yp is None:
            yp = torch.empty(0, device=x.device, dtype=y.dtype)
        ctx.mark_non_differentiable(r, yp)
        return y, r, yp

    @staticmethod
    def backward(ctx, dA, _dr, _dp):
        if dA is None:
            return (None,) * 9
        x, w, b = ctx.saved_tensors
        dA = dA.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            link = ctx.prev_link
            fused = getattr(K(), "conv_dgrad_actbwd", None)
            if link is not None and fused is not None and FUSE_ACT_BWD and not torch.is_grad_enabled():
                # plain backward: this data-gradient kernel also applies the backward of the
                # activation in front (and its bias gradient), see ActLink
                want_b = _bias_wanted(link.bias)
                tgt_b = _direct_target(link.bias) if want_b else None
                if not want_b or tgt_b is not None:
                    dx = fused(dA, w, ctx.op.adjoint(), ctx.scale, link.A, link.r if link.use_pn else None,
                               link.slope, link.use_pn, tgt_b)
                    if dx is not None:
                        link.fused = True
            if dx is None:
                dx = ConvFwd.apply(dA, w, ctx.op.adjoint(), ctx.scale)
        if _wants_grad(ctx, 1, w):
            tgt = _direct_target(w)
            if tgt is not None:
                K().conv_wgrad(x, dA, tuple(w.shape), ctx.op, ctx.scale, out=tgt)
            else:
                dw = ConvWgrad.apply(x, dA, ctx.op, ctx.scale, tuple(w.shape))
        if b is not None and _wants_grad(ctx, 2, b):
            if _direct_target(b) is not None:
                pass          # direct mode: accumulated by the Act/ActBwd kernels that produced dA
            else:
                db = ColSum.apply(dA)
        return dx, dw, db, None, None, None, None, None, None


def _bias_wanted(bias):
    if bias is None or not bias.requires_grad:
        return False
    if bias.is_leaf:
        try:
            return bool(torch._C._will_engine_execute_node(_acc_node(bias)))
        except Exception:
            return True
    return True


class Act(Function):
    """Graph marker turning the 'pre-activation handle' A into the activation y (same data).
    pool=True additionally applies the 2x2 average pool (bilinear x0.5 of the reference,
    progan_modules.py:299) so that its backward can be fused with the activation backward.
    `bias` is the bias parameter of the conv that produced A: every gradient that reaches A
    comes from this node's backward chain (ActBwd.forward and ActBwd.backward), so the bias
    gradient (column sum of dA) is produced there, fused into those kernels."""

    @staticmethod
    def forward(ctx, A, r, slope, use_pn, pool, bias, link=None, pooled=None):
        ctx.slope, ctx.use_pn, ctx.pool, ctx.bias = slope, use_pn, pool, bias
        ctx.link = link
        ctx.stash = {}
        ctx.save_for_backward(A, r)
        if pool:
            if pooled is not None and pooled.numel() > 0:
                return pooled.view_as(pooled)        # written by the conv's own epilogue
            return K().avgpool2(A, "nhwc")
        return A.view_as(A)

    @staticmethod
    def backward(ctx, dy):
        link = ctx.link
        if link is not None and link.fused:
            link.fused = False         # dy already is da: the consumer's kernel did our work
            return dy, None, None, None, None, None, None, None
        A, r = ctx.saved_tensors
        direct = DIRECT_GRADS and not torch.is_grad_enabled()
        addend = None
        ent = ctx.stash.pop("cot_a", None)
        if ent is not None:
            # the PixelNorm Hessian term left by ActBwd.backward earlier in THIS sweep: summed
            # into da (and into the bias gradient) by the kernel instead of by autograd
            if ent[1] != torch._C._current_graph_task_id():
                raise RuntimeError("progan_b200: stale second-order term (an earlier backward "
                                   "sweep did not reach this activation)")
            addend = ent[0]
        da = ActBwd.apply(dy.contiguous(), A, r, ctx.slope, ctx.use_pn, ctx.pool, ctx.bias, direct, ctx,
                          addend)
        return da, None, None, None, None, None, None, None


class ActBwd(Function):
    """da = Jpn(a)^T (m * dy)  — first-order PixelNorm+LeakyReLU backward (with the average
    pool's backward fused when pool=True) + the per-channel sum of da (bias gradient)."""

    @staticmethod
    def forward(ctx, dy, A, r, slope, use_pn, pool, bias, direct, act_node=None, addend=None):
        ctx.slope, ctx.use_pn, ctx.pool, ctx.bias = slope, use_pn, pool, bias
        ctx.act_node = act_node
        ctx.save_for_backward(dy, A, r)
        tgt = None
        want = _bias_wanted(bias)
        if direct and want and bias.is_leaf and bias.grad is not None:
            tgt = bias.grad
        if addend is not None:
            da, _ = K().pn_lrelu_bwd(dy, A, r if use_pn else None, slope, use_pn, pool, False, tgt, addend)
        else:
            da, _ = K().pn_lrelu_bwd(dy, A, r if use_pn else None, slope, use_pn, pool, False, tgt)
        return da

    @staticmethod
    @once_differentiable
    def backward(ctx, t):
        dy, A, r = ctx.saved_tensors
        cot_dy, cot_a = K().pn_lrelu_bwd_bwd(t.contiguous(), dy, A, r if ctx.use_pn else None,
                                             ctx.slope, ctx.use_pn, ctx.pool)
        if ctx.pool:
            cot_dy = K().avgpool2(cot_dy, "nhwc")
        if not ctx.needs_input_grad[1] or not ctx.use_pn:
            cot_a = None
        elif DIRECT_GRADS:
            node = ctx.act_node
            fus = getattr(K(), "actbwd_fusable", None)       # pass-through Act: no stash
            fusable = node is not None and node.link is not None and fus is not None and fus(A)
            if (STASH_SECOND_ORDER and node is not None and not fusable
                    and torch._C._will_engine_execute_node(node)):
                # the Act node of this layer runs later in this sweep (it receives the chain
                # term through the layer above): hand it the Hessian term directly — no autograd
                # add kernel, and its bias-gradient share is summed by that kernel too
                node.stash["cot_a"] = (cot_a, torch._C._current_graph_task_id())
                cot_a = None
            elif _bias_wanted(ctx.bias) and _direct_target(ctx.bias) is not None:
                K().colsum(cot_a, out=ctx.bias.grad)    # Hessian-term share of the bias gradient
        return cot_dy, cot_a, None, None, None, None, None, None, None, None


STASH_SECOND_ORDER = True


def conv_act(x, w, b, op, scale, slope=0.2, use_pn=True, pool=False, prev_link=None, make_link=False):
    """conv + bias + [PixelNorm] + LeakyReLU (+ 2x2 average pool).  make_link: also return the
    ActLink for the single conv that will consume the result (pass it there as prev_link)."""
    A, r, pooled = ConvAct.apply(x, w, b, op, scale, slope, use_pn, prev_link, pool)
    link = ActLink(A, r, slope, use_pn, b) if (make_link and not pool) else None
    y = Act.apply(A, r, slope, use_pn, pool, b, link, pooled)
    return (y, link) if make_link else y


# ---------------------------------------------------------------------- 1x1 heads
class PwFwd(Function):
    """from_rgb ('expand': image NCHW fp32 -> act NHWC) / to_rgb, linear ('reduce')."""

    @staticmethod
    def forward(ctx, x, w, b, kind, C, Kc, w_sc, w_sk, scale, act_dtype):
        ctx.cfg = (kind, C, Kc, w_sc, w_sk, scale, act_dtype)
        ctx.save_for_backward(x, w, b)
        if kind == "expand":
            return K().pw_expand(x, w, b, C, w_sc, w_sk, scale, act_dtype)
        return K().pw_reduce(x, w, b, Kc, w_sc, w_sk, scale)

    @staticmethod
    def backward(ctx, dy):
        kind, C, Kc, w_sc, w_sk, scale, act_dtype = ctx.cfg
        x, w, b = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        other = "reduce" if kind == "expand" else "expand"
        if ctx.needs_input_grad[0]:
            dx = PwFwd.apply(dy, w, None, other, C, Kc, w_sc, w_sk, scale, act_dtype)
        bias_done = False
        if _wants_grad(ctx, 1, w):
            act, img = (dy, x) if kind == "expand" else (x, dy)
            tgt = _direct_target(w)
            if tgt is not None:
                # from_rgb: the bias gradient (column sum of dy) rides on the same pass over dy
                tgt_b = None
                if kind == "expand" and b is not None and _wants_grad(ctx, 2, b):
                    tgt_b = _direct_target(b)
                if tgt_b is not None:
                    K().pw_wgrad(act, img, tuple(w.shape), w_sc, w_sk, scale, out=tgt, bias_out=tgt_b)
                    bias_done = True
                else:
                    K().pw_wgrad(act, img, tuple(w.shape), w_sc, w_sk, scale, out=tgt)
            else:
                dw = PwWgrad.apply(act, img, tuple(w.shape), C, Kc, w_sc, w_sk, scale, act_dtype)
        if b is not None and not bias_done and _wants_grad(ctx, 2, b):
            tgt = _direct_target(b)
            if tgt is not None:
                (K().colsum if kind == "expand" else K().img_chansum)(dy, out=tgt)
            else:
                db = ColSum.apply(dy) if kind == "expand" else ImgChanSum.apply(dy)
        return (dx, dw, db) + (None,) * 7


class PwWgrad(Function):
    @staticmethod
    def forward(ctx, act, img, wshape, C, Kc, w_sc, w_sk, scale, act_dtype):
        ctx.cfg = (C, Kc, w_sc, w_sk, scale, act_dtype)
        ctx.save_for_backward(act, img)
        return K().pw_wgrad(act, img, wshape, w_sc, w_sk, scale)

    @staticmethod
    def backward(ctx, ddw):
        C, Kc, w_sc, w_sk, scale, act_dtype = ctx.cfg
        act, img = ctx.saved_tensors
        ddw = ddw.contiguous()
        ca = ci = None
        if ctx.needs_input_grad[0]:
            ca = PwFwd.apply(img, ddw, None, "expand", C, Kc, w_sc, w_sk, scale, act_dtype)
        if ctx.needs_input_grad[1]:
            ci = PwFwd.apply(act, ddw, None, "reduce", C, Kc, w_sc, w_sk, scale, act_dtype)
        return (ca, ci) + (None,) * 7


class ImgChanSum(Function):
    @staticmethod
    def forward(ctx, img):
        ctx.shape = img.shape
        return K().img_chansum(img)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return g.view(1, -1, 1, 1).expand(ctx.shape).contiguous()


# --------------------------------------------------------------------- resampling
_PARTNER = {"avgpool2": "avgpool2_bwd", "avgpool2_bwd": "avgpool2",
            "upsample2": "upsample2_bwd", "upsample2_bwd": "upsample2"}


class Linear1(Function):
    """A parameter-free linear map and its transpose (bilinear x2 / x0.5 resampling)."""

    @staticmethod
    def forward(ctx, x, name, fmt):
        ctx.name, ctx.fmt = name, fmt
        return getattr(K(), name)(x, fmt)

    @staticmethod
    def backward(ctx, dy):
        return Linear1.apply(dy.contiguous(), _PARTNER[ctx.name], ctx.fmt), None, None


def avgpool2(x, fmt="nhwc"):
    return Linear1.apply(x, "avgpool2", fmt)


def upsample2(x, fmt="nhwc"):
    return Linear1.apply(x, "upsample2", fmt)


class Scale(Function):
    """out = (c0 + c1*alpha) * x, alpha read from device memory."""

    @staticmethod
    def forward(ctx, x, c0, c1, alpha_dev):
        ctx.c = (c0, c1)
        ctx.alpha_dev = alpha_dev
        return K().scale(x, c0, c1, alpha_dev)

    @staticmethod
    def backward(ctx, dy):
        return Scale.apply(dy.contiguous(), ctx.c[0], ctx.c[1], ctx.alpha_dev), None, None, None


class Blend(Function):
    """out = (1-alpha)*a + alpha*b   (progan_modules.py:212, 305)."""

    @staticmethod
    def forward(ctx, a, b, alpha_dev):
        ctx.alpha_dev = alpha_dev
        return K().blend(a, b, alpha_dev)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        da = Scale.apply(dy, 1.0, -1.0, ctx.alpha_dev) if ctx.needs_input_grad[0] else None
        db = Scale.apply(dy, 0.0, 1.0, ctx.alpha_dev) if ctx.needs_input_grad[1] else None
        return da, db, None


class Tanh(Function):
    @staticmethod
    def forward(ctx, x):
        y = K().tanh_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return K().tanh_bwd(dy.contiguous(), y)


# -------------------------------------------------------------------------- mbstd
def _groups(n, group):
    """Batch slices that each see their own minibatch statistic.  group=None: the whole batch
    (the reference, progan_modules.py:289-293).  The Trainer runs D once on cat([real, fake]) with
    group = per-pass batch, which reproduces the reference's two separate D calls exactly."""
    if not group or group >= n:
        return [(0, n)]
    if n % group:
        raise RuntimeError("progan_b200: batch %d is not a multiple of mbstd_group %d" % (n, group))
    return [(i, i + group) for i in range(0, n, group)]


class Mbstd(Function):
    @staticmethod
    def forward(ctx, x, Cp, group=None):
        gs = _groups(x.shape[0], group)
        outs, stats = zip(*[K().mbstd_fwd(x[a:b], Cp) for a, b in gs])
        ctx.stats, ctx.gs = stats, gs
        ctx.save_for_backward(x)
        return outs[0] if len(gs) == 1 else torch.cat(outs)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return MbstdBwd.apply(dout.contiguous(), x, ctx.stats, ctx.gs), None, None


class MbstdBwd(Function):
    @staticmethod
    def forward(ctx, dout, x, stats, gs):
        ctx.stats, ctx.gs = stats, gs
        ctx.save_for_backward(dout, x)
        dx = [K().mbstd_bwd(dout[a:b], x[a:b], st) for (a, b), st in zip(gs, stats)]
        return dx[0] if len(gs) == 1 else torch.cat(dx)

    @staticmethod
    @once_differentiable
    def backward(ctx, t):
        dout, x = ctx.saved_tensors
        t = t.contiguous()
        res = [K().mbstd_bwd_bwd(t[a:b], dout[a:b], x[a:b], st) for (a, b), st in zip(ctx.gs, ctx.stats)]
        if len(res) == 1:
            return res[0][0], res[0][1], None, None
        return torch.cat([r_[0] for r_ in res]), torch.cat([r_[1] for r_ in res]), None, None


# ---------------------------------------------------------------- gradient penalty
class GradPenalty(Function):
    """gp = lambda * mean_n (||g_n||_2 - 1)^2   (train.py:148-150)."""

    @staticmethod
    def forward(ctx, g, lam):
        g = g.contiguous()
        gp, norms = K().gp_fwd(g, lam)
        ctx.lam = lam
        ctx.save_for_backward(g, norms)
        return gp

    @staticmethod
    @once_differentiable
    def backward(ctx, up):
        g, norms = ctx.saved_tensors
        return K().gp_bwd(g, norms, up.contiguous(), ctx.lam), None


def gradient_penalty(grad_x_hat, lam=10.0):
    return GradPenalty.apply(grad_x_hat, lam)
