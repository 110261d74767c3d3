"""Host shim over the C-ABI: shape/dtype/contiguity checks, output allocation, launch.

Mirrors what the reference's C++ wrappers do around their kernels
(ada/torch_utils/ops/bias_act.cpp:32-97: TORCH_CHECK list, empty_like, launch on the
current stream) — but in Python, with torch used only as the allocator / stream owner.

Tensor conventions
  act   : [N,H,W,C] contiguous, float32 (check mode) or bfloat16 (product mode)
  image : [N,K,H,W] contiguous float32 (the module boundary, as the train scripts pass it)
  params, param grads, per-pixel statistics: float32
"""
import ctypes
import os
import weakref
from dataclasses import dataclass

import torch

from . import _lib

EPI_LINEAR, EPI_PN_LRELU, EPI_LRELU = 0, 1, 2
WL_TAP_CI_CO, WL_CO_TAP_CI, WL_TAP_CO_CI = 0, 1, 2

_DT = {torch.float32: 0, torch.bfloat16: 1}


@dataclass(frozen=True)
class ConvOp:
    """A 'logical' stride-1 cross-correlation y = conv_k,pad(x; Wl(w)).

    Wl is derived from the stored parameter w[d0,d1,k,k] by (swap, flip):
      swap=False: Wl[co][ci] = w[co][ci]   (nn.Conv2d, progan_modules.py:67)
      swap=True : Wl[co][ci] = w[ci][co]   (nn.ConvTranspose2d IOHW, :81, and data-grads)
      flip      : taps reversed.
    adjoint() is the op whose forward is this op's data-gradient.
    """
    k: int
    pad: int
    swap: bool = False
    flip: bool = False
    xpad: int = 0      # physical channel count of x when zero-padded beyond Cin (0 = not padded)
    ypad: int = 0      # physical channel count of y when zero-padded beyond Cout

    def adjoint(self):
        return ConvOp(self.k, self.k - 1 - self.pad, not self.swap, not self.flip, self.ypad, self.xpad)

    def cout(self, wshape):
        return wshape[1] if self.swap else wshape[0]

    def cin(self, wshape):
        return wshape[0] if self.swap else wshape[1]

    def cin_phys(self, wshape):
        return self.xpad or self.cin(wshape)

    def cout_phys(self, wshape):
        return self.ypad or self.cout(wshape)


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError("progan_b200: unsupported activation dtype %s" % t.dtype)


def _chk(t, name, dtype=None, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("progan_b200: %s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError(
            "progan_b200: %s is on %s; the kernels are CUDA-only (sm_100a) and there is no "
            "CPU fallback" % (name, t.device))
    if not t.is_contiguous():
        raise RuntimeError("progan_b200: %s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("progan_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError("progan_b200: %s must have %d dims, got %s" % (name, ndim, tuple(t.shape)))
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


class CudaKernels:
    """The product backend: every method is one (or two) launches of a hand-written kernel."""

    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        self.conv_impl = "tc"          # "tc": tcgen05 where the shape allows; "simt": always CUDA cores
        self.wgrad_tc = True           # tcgen05 weight-gradient kernels (False: CUDA-core wgrad)
        self.wide_tc = True            # 3x3 convs with 512/768/1024 output channels on tcgen05 (N tiles of 256)
        self.launches = 0              # kernels launched through this shim (bench `gpu_launches`)
        self._packs = {}               # id(param) -> [weakref, version, {variant: (tensor, pack args)}]
        self._pack_tables = {}         # ids of a parameter set -> (signature, device table, n)
        # data-gradient convs whose epilogue also applies the PixelNorm/LeakyReLU backward of the
        # layer in front (pg_conv_tc_actbwd: one thread per pixel row in the conv4 epilogue, 32 / 64
        # / 128 channels at >= 16 px).  Correct and tested, but OFF by default: measured on the
        # same box (profiles/r2/README.md) the step time is the same with and without it - the
        # stand-alone activation-backward kernels overlap the side-stream weight gradients, the
        # longer conv epilogue does not - while the conv kernel's FLOP rate drops from 947 to 840
        # TFLOP/s.  PG_FUSE_ACTBWD_MAX=128 (or setting fuse_actbwd_max_cout) turns it on.
        self.fuse_actbwd_min_cout = int(os.environ.get("PG_FUSE_ACTBWD_MIN", "32"))
        self.fuse_actbwd_max_cout = int(os.environ.get("PG_FUSE_ACTBWD_MAX", "0"))
        self.wgrad_side_stream = None  # Trainer: deferred weight gradients run on this stream, next
        self._side_dirty = set()       # to the bandwidth-bound kernels of the data-gradient chain
        self._side_streams = {}        # one side stream per forking stream
        self.defer_wgrad = False       # Trainer: weight gradients accumulate in persistent workspaces
        self._wgrad_ws = {}            # (grad ptr, variant) -> (workspace, unpack entry)
        self._pending = {}             # workspaces holding partial sums since the last flush
        self._unpack_tables = {}

    # ------------------------------------------------------------------ utils
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            _lib.check(rc, name)
        self.launches += 1

    def tc_mode(self, dtype, H, W, wshape, op):
        """Which tcgen05 form serves this conv: 'conv3' (3x3 pad 1 implicit GEMM), 'valid'
        (kxk valid conv of a kxk map == GEMM with K = k*k*Cin), 'full' (kxk full conv of a 1x1
        map == GEMM with N = k*k*Cout), or None (SIMT kernels)."""
        if self.conv_impl != "tc" or dtype != torch.bfloat16:
            return None
        cin, cout = op.cin_phys(wshape), op.cout_phys(wshape)
        if cin % 32 or cout % 32:
            return None
        def fits(c):            # above 256 channels: N tiles of 256, PixelNorm as its own kernel
            return c <= 256 or (self.wide_tc and c % 256 == 0 and c <= 1024)

        if op.k == 3 and op.pad == 1 and fits(cout):
            return "conv3"
        if op.xpad or op.ypad:
            return None
        if op.pad == 0 and H == op.k and W == op.k and fits(cout):
            return "valid"
        if op.pad == op.k - 1 and H == 1 and W == 1 and fits(cout) and fits(cin):
            return "full"
        return None

    def tc_eligible(self, x, wshape, op):
        return self.tc_mode(x.dtype, x.shape[1], x.shape[2], wshape, op) is not None

    # --------------------------------------------------------------- packing
    def invalidate_packs(self):
        self._packs.clear()
        self._pack_tables.clear()

    def drop_packs(self, params):
        """Forget the operand copies of these parameters (they were modified behind autograd's
        back, e.g. by the EMA kernel, and are not worth refreshing eagerly)."""
        for p_ in params:
            self._packs.pop(id(p_), None)

    @staticmethod
    def _upload(entries, cls, device):
        arr = (cls * len(entries))(*entries)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        return raw.to(device)

    def packed(self, w, op, layout, dtype, flip=None):
        """Operand-layout copy of weight_orig for (op, layout, dtype); cached per parameter
        version (in-place optimizer updates bump Tensor._version)."""
        flip = op.flip if flip is None else flip
        key = (op.swap, flip, layout, dtype, op.xpad, op.ypad)
        cacheable = isinstance(w, torch.nn.Parameter)
        if cacheable:
            ent = self._packs.get(id(w))
            if ent is not None and ent[0]() is w:
                if ent[1] != w._version:
                    # the parameter changed behind the cache (load_state_dict, an in-place edit):
                    # re-pack every cached copy INTO ITS OWN STORAGE - captured CUDA graphs and
                    # the device pack tables hold these pointers
                    for out_, a_ in ent[2].values():
                        self._call("pg_pack_conv_weight", w.data_ptr(), out_.data_ptr(), *a_, self._stream())
                    ent[1] = w._version
                hit = ent[2].get(key)
                if hit is not None:
                    return hit[0]
            else:
                ent = [weakref.ref(w), w._version, {}]
                self._packs[id(w)] = ent
        _chk(w, "weight", torch.float32, 4)
        d0, d1, kh, kw = w.shape
        cout, cin = op.cout_phys(w.shape), op.cin_phys(w.shape)
        out = torch.empty(cout * kh * kw * cin, device=w.device, dtype=dtype)
        args = (d0, d1, kh, kw, int(op.swap), int(flip), layout, cin, cout, _DT[dtype])
        self._call("pg_pack_conv_weight", w.data_ptr(), out.data_ptr(), *args, self._stream())
        if cacheable:
            ent[2][key] = (out, args)
        return out

    def refresh_packs(self, params):
        """Re-pack every cached operand copy of `params` in ONE launch (after an optimizer step
        that wrote the parameters through the flat bucket).  The copies keep their storage, so a
        captured CUDA graph stays valid."""
        ents = []
        for p_ in params:
            ent = self._packs.get(id(p_))
            if ent is not None and ent[0]() is p_:
                ents.append((p_, ent))
        sig = tuple((p_.data_ptr(), v[0].data_ptr(), k) for p_, ent in ents for k, v in ent[2].items())
        if not sig:
            return
        tkey = tuple(id(p_) for p_ in params)
        tab = self._pack_tables.get(tkey)
        if tab is None or tab[0] != sig:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("progan_b200: pack table changed during CUDA-graph capture "
                                   "(run warm-up iterations before capturing)")
            rows = []
            for p_, ent in ents:
                for k, (out, a) in ent[2].items():
                    d0, d1, kh, kw, swap, flip, layout, cin, cout, dt = a
                    rows.append(_lib.PackEntry(p_.data_ptr(), out.data_ptr(), cout * kh * kw * cin, d0, d1,
                                               kh * kw, swap, flip, layout, cin, cout, dt, 0))
            tab = (sig, self._upload(rows, _lib.PackEntry, params[0].device), len(rows))
            self._pack_tables[tkey] = tab
        self._call("pg_pack_conv_weight_multi", tab[1].data_ptr(), tab[2], self._stream())
        for p_, ent in ents:
            ent[1] = p_._version

    # ------------------------------------------------------------------ conv
    def conv_fwd(self, x, w, bias, op, scale, epi=EPI_LINEAR, slope=0.2, pool_out=False):
        """y = epi(scale * conv(x; Wl(w)) + bias).  Returns (y, r) with r the per-pixel
        PixelNorm rsqrt (fp32 [N,Ho,Wo]) for EPI_PN_LRELU, else None.  pool_out: returns
        (y, r, y_pool) where y_pool = avgpool2(y) written by the same kernel, or None when the
        shape is not served by the fused epilogue (the caller then pools separately)."""
        _chk(x, "x", ndim=4)
        _chk(w, "w", torch.float32, 4)
        if bias is not None:
            _chk(bias, "bias", torch.float32, 1)
        N, H, W, C = x.shape
        k = op.k
        cin, cout = op.cin_phys(w.shape), op.cout_phys(w.shape)
        if C != cin or w.shape[2] != k or w.shape[3] != k:
            raise RuntimeError("progan_b200: conv shape mismatch x=%s w=%s op=%s"
                               % (tuple(x.shape), tuple(w.shape), op))
        Ho, Wo = H + 2 * op.pad - k + 1, W + 2 * op.pad - k + 1
        y = torch.empty((N, Ho, Wo, cout), device=x.device, dtype=x.dtype)
        r = torch.empty((N, Ho, Wo), device=x.device, dtype=torch.float32) if epi == EPI_PN_LRELU else None
        mode = self.tc_mode(x.dtype, H, W, w.shape, op)
        nb = bias.numel() if bias is not None else 0
        st = self._stream()
        yp = None
        if mode == "conv3":
            wp = self.packed(w, op, WL_CO_TAP_CI, torch.bfloat16)
            if pool_out and H % 16 == 0 and W % 8 == 0 and cout in (32, 64, 128) and not (op.xpad or op.ypad):
                yp = torch.empty((N, H // 2, W // 2, cout), device=x.device, dtype=x.dtype)
            if cout > 256:
                # wide layer (the Correct* defaults of 512): one pixel's channel vector spans several
                # N tiles, so the conv writes the pre-activation and PixelNorm runs stand-alone
                conv_epi = epi if epi != EPI_PN_LRELU else EPI_LINEAR
                self._call("pg_conv_tc", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(), None,
                           N, H, W, cin, cout, 256, 9, nb, float(scale), conv_epi, float(slope), None, st)
                if epi == EPI_PN_LRELU:
                    self.pn_lrelu_fwd(y, slope, True, out=y, r=r)
            else:
                self._call("pg_conv_tc", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(), _ptr(r),
                           N, H, W, cin, cout, cout, 9, nb, float(scale), epi, float(slope), _ptr(yp), st)
        elif mode in ("valid", "full"):
            wide = cout > 256
            conv_epi = EPI_LINEAR if (wide and epi == EPI_PN_LRELU) else epi
            tile = 256 if wide else cout
            if mode == "valid":    # [N, k*k*cin] x [cout, k*k*cin]^T
                wp = self.packed(w, op, WL_CO_TAP_CI, torch.bfloat16)
                self._call("pg_conv_tc", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(),
                           None if wide else _ptr(r), N, 1, 1, k * k * cin, cout, tile, 1, nb, float(scale),
                           conv_epi, float(slope), None, st)
            else:                  # [N, cin] x [(pos, cout), cin]^T ; output position = flipped tap
                wp = self.packed(w, op, WL_TAP_CO_CI, torch.bfloat16, flip=not op.flip)
                self._call("pg_conv_tc", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(),
                           None if wide else _ptr(r), N, 1, 1, cin, k * k * cout, tile, 1, nb, float(scale),
                           conv_epi, float(slope), None, st)
            if wide and epi == EPI_PN_LRELU:
                self.pn_lrelu_fwd(y, slope, True, out=y, r=r)
        else:
            wp = self.packed(w, op, WL_TAP_CI_CO, x.dtype)
            self._call("pg_conv_fwd_simt", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(),
                       _ptr(r), N, H, W, cin, cout, k, op.pad, float(scale), epi, float(slope),
                       _dt(x), st)
        return (y, r, yp) if pool_out else (y, r)

    def conv_dgrad_actbwd(self, x, w, op, scale, y_prev, r_prev, slope, use_pn, colsum_out=None):
        """Fused da_prev = Jpn(a_prev)^T (m * (scale * conv(x; Wl(w)))): a data-gradient conv whose
        epilogue applies the PixelNorm+LeakyReLU backward of the layer in front (stored activation
        y_prev, statistic r_prev) and adds the per-channel sum of da_prev to colsum_out (that
        layer's bias gradient).  Returns None when the fused kernel does not take the shape."""
        if x.dtype != torch.bfloat16 or op.xpad or op.ypad:
            return None
        N, H, W, C = x.shape
        if self.tc_mode(x.dtype, H, W, w.shape, op) != "conv3":
            return None
        cin, cout = op.cin_phys(w.shape), op.cout_phys(w.shape)
        if C != cin or tuple(y_prev.shape) != (N, H, W, cout) or not self.actbwd_fusable(y_prev):
            return None
        _chk(x, "x", torch.bfloat16, 4)
        _chk(y_prev, "y_prev", torch.bfloat16, 4)
        if use_pn:
            _chk(r_prev, "r_prev", torch.float32)
        if colsum_out is not None:
            _chk(colsum_out, "colsum_out", torch.float32, 1)
        wp = self.packed(w, op, WL_CO_TAP_CI, torch.bfloat16)
        da = torch.empty((N, H, W, cout), device=x.device, dtype=x.dtype)
        try:
            self._call("pg_conv_tc_actbwd", x.data_ptr(), wp.data_ptr(), da.data_ptr(), N, H, W, cin, cout,
                       float(scale), y_prev.data_ptr(), _ptr(r_prev) if use_pn else None,
                       float(slope), int(use_pn), _ptr(colsum_out), self._stream())
        except RuntimeError as e:
            if "(-3)" in str(e):       # PG_ERR_UNSUPPORTED: caller runs the two-kernel path
                return None
            raise
        return da

    def actbwd_fusable(self, y):
        """Is the backward of this stored activation [N,H,W,C] served by the fused data-gradient
        epilogue (pg_conv_tc_actbwd)?"""
        if self.conv_impl != "tc" or y.dtype != torch.bfloat16 or y.dim() != 4:
            return False
        _, H, W, C = y.shape
        return (C in (32, 64, 128) and self.fuse_actbwd_min_cout <= C <= self.fuse_actbwd_max_cout
                and H % 16 == 0 and W % 8 == 0)

    def conv_wgrad(self, x, dy, wshape, op, scale, out=None):
        """dw[wshape] = scale * sum_pix dy (x) x for the conv `op` (fp32).  out: accumulate into
        this fp32 tensor (a parameter's .grad view of the flat bucket) instead of allocating."""
        _chk(x, "x", ndim=4)
        _chk(dy, "dy", x.dtype, 4)
        N, H, W, cin = x.shape
        cout = dy.shape[-1]
        k = op.k
        if cin != op.cin_phys(wshape) or cout != op.cout_phys(wshape):
            raise RuntimeError("progan_b200: wgrad shape mismatch")
        cin_l, cout_l = op.cin(wshape), op.cout(wshape)
        mode = self.tc_mode(x.dtype, H, W, wshape, op) if self.wgrad_tc else None
        st = self._stream()
        acc = 0
        if out is not None:
            _chk(out, "out", torch.float32)
            if tuple(out.shape) != tuple(wshape):
                raise RuntimeError("progan_b200: wgrad out shape mismatch")
            acc = 1
        if mode is not None:
            dw = out if out is not None else torch.empty(tuple(wshape), device=x.device, dtype=torch.float32)
            if out is not None and self.defer_wgrad:
                # accumulate into the parameter's persistent workspace; flush_wgrads() folds every
                # pending workspace into its gradient with one launch
                wkey = (out.data_ptr(), mode, op.swap, op.flip, cin, cout, cin_l, cout_l, k, float(scale))
                went = self._wgrad_ws.get(wkey)
                if went is None:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("progan_b200: new weight-gradient workspace during CUDA-graph "
                                           "capture (run warm-up iterations before capturing)")
                    ws = torch.zeros(k * k * cin * cout, device=x.device, dtype=torch.float32)
                    if mode == "conv3":
                        u = (cin_l, cout_l, cin, cout, 9, int(op.swap), int(op.flip))
                    elif mode == "valid":
                        u = (cin, cout, cin, cout, k * k, int(op.swap), int(op.flip))
                    else:
                        u = (cout, cin, cout, cin, k * k, int(not op.swap), int(not op.flip))
                    went = (ws, _lib.UnpackEntry(ws.data_ptr(), out.data_ptr(), *u, 0, float(scale), 0.0))
                    self._wgrad_ws[wkey] = went
                ws = went[0]
                self._pending[wkey] = went
                acc = 2
            else:
                ws = torch.empty(k * k * cin * cout, device=x.device, dtype=torch.float32)
            side = self._side_stream_for_current() if acc == 2 else None
            if side is not None:
                # fork: the weight gradient only needs x and dy, nothing on the main stream needs
                # its result before flush_wgrads() — let it overlap the activation-backward /
                # resampling kernels that follow (tensor-bound next to HBM-bound)
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                x.record_stream(side)
                dy.record_stream(side)
                self._side_dirty.add(side)
            with (torch.cuda.stream(side) if side is not None else _nullctx()):
                st = self._stream()
                if mode == "conv3":
                    self._call("pg_conv_wgrad_tc", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                               N, H, W, cin, cout, cin_l, cout_l, 9, 0, float(scale), int(op.swap),
                               int(op.flip), acc, st)
                elif mode == "valid":
                    self._call("pg_conv_wgrad_tc", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                               N, 1, 1, cin, cout, cin, cout, k * k, 1, float(scale), int(op.swap),
                               int(op.flip), acc, st)
                else:   # "full": the same quantity as the valid-form weight gradient of the adjoint op
                    self._call("pg_conv_wgrad_tc", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                               N, 1, 1, cout, cin, cout, cin, k * k, 1, float(scale), int(not op.swap),
                               int(not op.flip), acc, st)
            if acc != 2:
                self.launches += 2          # memset + unpack
        else:
            if op.xpad or op.ypad:
                raise RuntimeError("progan_b200: padded channels are only used on the tcgen05 path")
            dw = out if out is not None else torch.zeros(tuple(wshape), device=x.device, dtype=torch.float32)
            self._call("pg_conv_wgrad_simt", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W,
                       cin, cout, k, op.pad, float(scale), int(op.swap), int(op.flip), _dt(x), st)
        return dw

    def _side_stream_for_current(self):
        """The weight-gradient stream paired with the current stream (None when the feature is
        off): concurrent passes on different streams each get their own."""
        if self.wgrad_side_stream is None:
            return None
        cur = torch.cuda.current_stream()
        side = self._side_streams.get(cur.cuda_stream)
        if side is None:
            side = self.wgrad_side_stream if not self._side_streams else torch.cuda.Stream()
            self._side_streams[cur.cuda_stream] = side
        return side

    def flush_wgrads(self, ptr_range=None, join=True):
        """Fold pending weight-gradient workspaces into their gradients: one launch per
        'round', where a round holds at most one workspace per gradient tensor (a conv's own
        weight gradient and its adjoint form from the GP sweep go to different rounds, so the
        kernel's read-modify-write of dw needs no atomics).  ptr_range = (lo, hi): only the
        workspaces whose gradient lives in that address range (the data-parallel Trainer flushes
        and all-reduces the top of the critic while the backward sweep is still running); join:
        wait for the weight-gradient side streams first (the caller did it itself otherwise)."""
        if not self._pending:
            return
        if ptr_range is None:
            sel = dict(self._pending)
        else:
            sel = {k: v for k, v in self._pending.items() if ptr_range[0] <= k[0] < ptr_range[1]}
            if not sel:
                return
        if join:
            for side in self._side_dirty:    # join the weight-gradient streams
                torch.cuda.current_stream().wait_stream(side)
            self._side_dirty = set()
        sig = tuple(sel.keys())
        tabs = self._unpack_tables.get(sig)
        if tabs is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("progan_b200: unpack table changed during CUDA-graph capture")
            rounds = []
            for wkey, (ws, ent) in sel.items():
                for r in rounds:
                    if wkey[0] not in r[0]:
                        break
                else:
                    r = (set(), [])
                    rounds.append(r)
                r[0].add(wkey[0])
                r[1].append(ent)
            dev = next(iter(sel.values()))[0].device
            tabs = [(self._upload(rows, _lib.UnpackEntry, dev), len(rows)) for _, rows in rounds]
            self._unpack_tables[sig] = tabs
        for tab, n in tabs:
            self._call("pg_wgrad_unpack_multi", tab.data_ptr(), n, self._stream())
        for k in sel:
            del self._pending[k]

    def mbstd_channels(self, C, dtype):
        """Physical channel count of the minibatch-stddev output (C real + 1 statistic):
        padded to a multiple of 32 on the tensor-core path."""
        if self.conv_impl == "tc" and dtype == torch.bfloat16 and C % 32 == 0:
            Cp = ((C + 1 + 31) // 32) * 32
            if Cp <= 256:
                return Cp
            if self.wide_tc and C <= 768:      # wide layers tile the channels in blocks of 256
                return ((C + 1 + 255) // 256) * 256
        return C + 1

    # ------------------------------------------------- PixelNorm + LeakyReLU
    def pn_lrelu_fwd(self, a, slope, use_pn, out=None, r=None):
        """(y, r): y = lrelu(a * r), r = rsqrt(mean_c a^2 + 1e-8) per pixel (None without PixelNorm).
        The stand-alone form for layers too wide for the conv epilogue; out may alias a."""
        _chk(a, "a", ndim=4)
        C = a.shape[-1]
        y = out if out is not None else torch.empty_like(a)
        if use_pn and r is None:
            r = torch.empty(a.shape[:-1], device=a.device, dtype=torch.float32)
        self._call("pg_pn_lrelu_fwd", a.data_ptr(), y.data_ptr(), _ptr(r) if use_pn else None,
                   a.numel() // C, C, float(slope), int(use_pn), _dt(a), self._stream())
        return y, (r if use_pn else None)

    def pn_lrelu_bwd(self, dy, y, r, slope, use_pn, pool=False, want_colsum=False, colsum_out=None,
                     addend=None):
        """da = Jpn(a)^T (m*dy) [+ addend].  pool: dy is the gradient of avgpool2(y).  Returns
        (da, colsum) where colsum = per-channel sum of da (bias gradient) when requested, else None."""
        if addend is not None:
            _chk(addend, "addend", y.dtype)
            if addend.shape != y.shape:
                raise RuntimeError("progan_b200: addend shape mismatch")
        _chk(dy, "dy", y.dtype)
        _chk(y, "y")
        N, H, W, C = y.shape
        if pool and tuple(dy.shape) != (N, H // 2, W // 2, C):
            raise RuntimeError("progan_b200: pooled dy shape mismatch")
        da = torch.empty_like(y)
        cs = colsum_out
        if cs is None and want_colsum:
            cs = torch.zeros(C, device=y.device, dtype=torch.float32)
        self._call("pg_pn_lrelu_bwd", dy.data_ptr(), y.data_ptr(), _ptr(r), da.data_ptr(),
                   y.numel() // C, C, float(slope), int(use_pn), H if pool else 0, W if pool else 0,
                   _ptr(cs), _ptr(addend), _dt(y), self._stream())
        return da, cs

    def pn_lrelu_bwd_bwd(self, t, dy, y, r, slope, use_pn, pool=False):
        """(cot_dy at full resolution, cot_a); pool: dy is the pooled-activation gradient."""
        _chk(t, "t", y.dtype)
        _chk(dy, "dy", y.dtype)
        N, H, W, C = y.shape
        cot_dy, cot_a = torch.empty_like(y), torch.empty_like(y)
        self._call("pg_pn_lrelu_bwd_bwd", t.data_ptr(), dy.data_ptr(), y.data_ptr(), _ptr(r),
                   cot_dy.data_ptr(), cot_a.data_ptr(), y.numel() // C, C, float(slope),
                   int(use_pn), H if pool else 0, W if pool else 0, _dt(y), self._stream())
        return cot_dy, cot_a

    def colsum(self, x, out=None):
        _chk(x, "x")
        C = x.shape[-1]
        if out is None:
            out = torch.zeros(C, device=x.device, dtype=torch.float32)
        self._call("pg_colsum", x.data_ptr(), out.data_ptr(), x.numel() // C, C, _dt(x), self._stream())
        return out

    # ------------------------------------------------------------ 1x1 heads
    def pw_expand(self, img, w, bias, C, w_sc, w_sk, scale, dtype):
        _chk(img, "img", torch.float32, 4)
        _chk(w, "w", torch.float32)
        N, Kc, H, W = img.shape
        act = torch.empty((N, H, W, C), device=img.device, dtype=dtype)
        self._call("pg_pw_expand", img.data_ptr(), w.data_ptr(), _ptr(bias), act.data_ptr(), N, H * W,
                   Kc, C, w_sc, w_sk, float(scale), _DT[dtype], self._stream())
        return act

    def pw_reduce(self, act, w, bias, Kc, w_sc, w_sk, scale):
        _chk(act, "act", ndim=4)
        _chk(w, "w", torch.float32)
        N, H, W, C = act.shape
        img = torch.empty((N, Kc, H, W), device=act.device, dtype=torch.float32)
        self._call("pg_pw_reduce", act.data_ptr(), w.data_ptr(), _ptr(bias), img.data_ptr(), N, H * W,
                   Kc, C, w_sc, w_sk, float(scale), _dt(act), self._stream())
        return img

    def pw_wgrad(self, act, img, wshape, w_sc, w_sk, scale, out=None, bias_out=None):
        """dw (+= into `out`); bias_out: fp32 [C] that additionally receives the per-channel sum of
        `act` (the bias gradient of a from_rgb layer) from the same pass."""
        _chk(act, "act", ndim=4)
        _chk(img, "img", torch.float32, 4)
        N, H, W, C = act.shape
        Kc = img.shape[1]
        if bias_out is not None:
            _chk(bias_out, "bias_out", torch.float32, 1)
            if bias_out.numel() != C:
                raise RuntimeError("progan_b200: pw_wgrad bias_out must have %d elements" % C)
        dw = out if out is not None else torch.zeros(tuple(wshape), device=act.device, dtype=torch.float32)
        self._call("pg_pw_wgrad", act.data_ptr(), img.data_ptr(), dw.data_ptr(), _ptr(bias_out), N, H * W,
                   Kc, C, w_sc, w_sk, float(scale), _dt(act), self._stream())
        return dw

    def img_chansum(self, img, out=None):
        _chk(img, "img", torch.float32, 4)
        N, Kc, H, W = img.shape
        if out is None:
            out = torch.zeros(Kc, device=img.device, dtype=torch.float32)
        self._call("pg_img_chansum", img.data_ptr(), out.data_ptr(), N, H * W, Kc, self._stream())
        return out

    # ------------------------------------------------------------- resample
    def _resample(self, fn, x, fmt, out_hw):
        _chk(x, "x", ndim=4)
        if fmt == "nchw":
            N, Kc, H, W = x.shape
            n_, c_ = N * Kc, 1
            out = torch.empty((N, Kc) + out_hw(H, W), device=x.device, dtype=x.dtype)
        else:
            N, H, W, C = x.shape
            n_, c_ = N, C
            out = torch.empty((N,) + out_hw(H, W) + (C,), device=x.device, dtype=x.dtype)
        return out, n_, H, W, c_

    def avgpool2(self, x, fmt="nhwc"):
        out, n, H, W, c = self._resample("avgpool2", x, fmt, lambda h, w: (h // 2, w // 2))
        self._call("pg_avgpool2", x.data_ptr(), out.data_ptr(), n, H, W, c, _dt(x), self._stream())
        return out

    def avgpool2_bwd(self, dy, fmt="nhwc"):
        out, n, H, W, c = self._resample("avgpool2_bwd", dy, fmt, lambda h, w: (2 * h, 2 * w))
        self._call("pg_avgpool2_bwd", dy.data_ptr(), out.data_ptr(), n, 2 * H, 2 * W, c, _dt(dy),
                   self._stream())
        return out

    def upsample2(self, x, fmt="nhwc"):
        out, n, H, W, c = self._resample("upsample2", x, fmt, lambda h, w: (2 * h, 2 * w))
        self._call("pg_upsample2", x.data_ptr(), out.data_ptr(), n, H, W, c, _dt(x), self._stream())
        return out

    def upsample2_bwd(self, dy, fmt="nhwc"):
        out, n, H, W, c = self._resample("upsample2_bwd", dy, fmt, lambda h, w: (h // 2, w // 2))
        self._call("pg_upsample2_bwd", dy.data_ptr(), out.data_ptr(), n, H // 2, W // 2, c, _dt(dy),
                   self._stream())
        return out

    # ---------------------------------------------------------- elementwise
    def blend(self, a, b, alpha_dev):
        _chk(a, "a")
        _chk(b, "b", a.dtype)
        _chk(alpha_dev, "alpha_dev", torch.float32)
        if a.shape != b.shape:
            raise RuntimeError("progan_b200: blend shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
        out = torch.empty_like(a)
        self._call("pg_blend", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(),
                   alpha_dev.data_ptr(), _dt(a), self._stream())
        return out

    def scale(self, x, c0, c1, alpha_dev):
        _chk(x, "x")
        out = torch.empty_like(x)
        self._call("pg_scale", x.data_ptr(), out.data_ptr(), x.numel(), float(c0), float(c1),
                   _ptr(alpha_dev), _dt(x), self._stream())
        return out

    def tanh_fwd(self, x):
        _chk(x, "x", torch.float32)
        y = torch.empty_like(x)
        self._call("pg_tanh_fwd", x.data_ptr(), y.data_ptr(), x.numel(), self._stream())
        return y

    def tanh_bwd(self, dy, y):
        _chk(dy, "dy", torch.float32)
        _chk(y, "y", torch.float32)
        dx = torch.empty_like(y)
        self._call("pg_tanh_bwd", dy.data_ptr(), y.data_ptr(), dx.data_ptr(), y.numel(), self._stream())
        return dx

    # --------------------------------------------------------------- mbstd
    def mbstd_fwd(self, x, Cp):
        """Returns (out, stats): stats is the fp32 workspace (mu, sigma per feature) reused by
        the backward kernels."""
        _chk(x, "x", ndim=4)
        N, H, W, C = x.shape
        if H != 4 or W != 4:
            raise RuntimeError("progan_b200: minibatch-stddev expects a 4x4 map, got %dx%d" % (H, W))
        out = torch.empty((N, 4, 4, Cp), device=x.device, dtype=x.dtype)
        stats = torch.empty(4 * 16 * C + 64, device=x.device, dtype=torch.float32)
        self._call("pg_mbstd_fwd", x.data_ptr(), out.data_ptr(), stats.data_ptr(), N, C, Cp, _dt(x),
                   self._stream())
        self.launches += 1
        return out, stats

    def mbstd_bwd(self, dout, x, stats):
        _chk(dout, "dout", x.dtype, 4)
        N, _, _, C = x.shape
        dx = torch.empty_like(x)
        self._call("pg_mbstd_bwd", dout.data_ptr(), x.data_ptr(), stats.data_ptr(), dx.data_ptr(), N,
                   C, dout.shape[-1], _dt(x), self._stream())
        return dx

    def mbstd_bwd_bwd(self, t, dout, x, stats):
        _chk(t, "t", x.dtype, 4)
        _chk(dout, "dout", x.dtype, 4)
        N, _, _, C = x.shape
        cot_dout, cot_x = torch.empty_like(dout), torch.empty_like(x)
        st2 = torch.empty_like(stats)       # (tbar, c_f) of this call; mu/sigma are recomputed
        self._call("pg_mbstd_bwd_bwd", t.data_ptr(), dout.data_ptr(), x.data_ptr(), st2.data_ptr(),
                   cot_dout.data_ptr(), cot_x.data_ptr(), N, C, dout.shape[-1], _dt(x), self._stream())
        self.launches += 1
        return cot_dout, cot_x

    # ------------------------------------------------------------- WGAN-GP
    def wgan_loss(self, d, n_real, drift, metric=None):
        """Gradient seed of the critic loss (n_real > 0: real outputs first, train.py:126-139) or
        of the generator loss (n_real = 0, :162-167) w.r.t. the critic outputs d; the logged loss
        value is added to the device scalar `metric`."""
        _chk(d, "d", torch.float32)
        n = d.numel()
        seed = torch.empty_like(d)
        if metric is not None:
            _chk(metric, "metric", torch.float32)
        self._call("pg_wgan_loss", d.data_ptr(), seed.data_ptr(), _ptr(metric), int(n_real), n - int(n_real),
                   float(drift), self._stream())
        return seed

    def interp_xhat(self, real, fake, eps):
        _chk(real, "real", torch.float32)
        _chk(fake, "fake", torch.float32)
        _chk(eps, "eps", torch.float32)
        N = real.shape[0]
        if fake.shape != real.shape or eps.numel() != N:
            raise RuntimeError("progan_b200: interp_xhat shape mismatch")
        out = torch.empty_like(real)
        self._call("pg_interp_xhat", real.data_ptr(), fake.data_ptr(), eps.data_ptr(), out.data_ptr(),
                   N, real.numel() // N, self._stream())
        return out

    def gp_fwd(self, g, lam):
        _chk(g, "g", torch.float32)
        N = g.shape[0]
        norms = torch.empty(N, device=g.device, dtype=torch.float32)
        gp = torch.empty((), device=g.device, dtype=torch.float32)
        self._call("pg_gp_fwd", g.data_ptr(), norms.data_ptr(), gp.data_ptr(), N, g.numel() // N,
                   float(lam), self._stream())
        self.launches += 1
        return gp, norms

    def gp_bwd(self, g, norms, upstream, lam):
        _chk(g, "g", torch.float32)
        N = g.shape[0]
        v = torch.empty_like(g)
        self._call("pg_gp_bwd", g.data_ptr(), norms.data_ptr(), _ptr(upstream), v.data_ptr(), N,
                   g.numel() // N, float(lam), self._stream())
        return v

    # ----------------------------------------------------------- optimiser
    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, step_dev, grad_scale=1.0):
        for t, n in ((p, "p"), (g, "g"), (v, "v"), (step_dev, "step")):
            _chk(t, n, torch.float32)
        self._call("pg_adam_step", p.data_ptr(), g.data_ptr(), _ptr(m), v.data_ptr(), p.numel(),
                   float(lr), float(beta1), float(beta2), float(eps), step_dev.data_ptr(),
                   float(grad_scale), self._stream())

    def adam_multi(self, p, g, m, v, chunks, steps_dev, lr, beta1, beta2, eps, grad_scale=1.0):
        for t, n in ((p, "p"), (g, "g"), (v, "v"), (steps_dev, "steps")):
            _chk(t, n, torch.float32)
        _chk(chunks, "chunks", torch.int32, 2)
        self._call("pg_adam_multi", p.data_ptr(), g.data_ptr(), _ptr(m), v.data_ptr(),
                   chunks.data_ptr(), chunks.shape[0], steps_dev.data_ptr(), float(lr), float(beta1),
                   float(beta2), float(eps), float(grad_scale), self._stream())

    def ema(self, ema, p, decay):
        _chk(ema, "ema", torch.float32)
        _chk(p, "p", torch.float32)
        self._call("pg_ema", ema.data_ptr(), p.data_ptr(), p.numel(), float(decay), self._stream())


_backend = None


def get_kernels():
    """The process-wide kernel backend (loads the shared library on first use)."""
    global _backend
    if _backend is None:
        _backend = CudaKernels()
    return _backend


def set_kernels(backend):
    """Install a different backend object (tests install a torch emulation of the kernel
    semantics to check the host/autograd logic on machines without a GPU)."""
    global _backend
    prev = _backend
    _backend = backend
    return prev
