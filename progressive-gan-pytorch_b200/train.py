"""The training hot loop body of the reference (train.py:97-169) on the B200 kernels.

`Trainer.step(real, z, eps, step, alpha)` performs exactly one iteration of the reference
loop with n_critic = 1:

    D.zero_grad; D(real) loss with 0.001 drift, backward          train.py:98,126-130
    fake = G(z); D(fake.detach()).mean().backward                 train.py:133-139
    x_hat = eps*real + (1-eps)*fake; grad(D(x_hat).sum(), x_hat, create_graph=True)
    gp = 10*mean((||g||-1)^2); gp.backward                        train.py:142-151
    d_optimizer.step()                                            train.py:155
    G.zero_grad; loss = -D(fake).mean(); backward; g_optimizer.step(); EMA   :158-169

What is different from the reference (all host-side, numerics unchanged):
  * parameters / gradients / Adam state of each net live in ONE flat fp32 bucket, ordered so
    the parameters that are live at a given (step, fading) form at most two contiguous
    ranges; Adam is one multi-tensor launch per net, EMA one launch, zero_grad one memset;
  * torch.optim.Adam semantics are kept per parameter group (own step counter, parameters
    without gradient untouched — the reference relies on `grad is None` skipping);
  * data parallel: one process per GPU, gradients of the live ranges are all-reduced (NCCL)
    and averaged inside the Adam kernel (grad_scale = 1/world); minibatch-stddev stays
    per-rank;
  * losses are accumulated on the device (no .item() sync per iteration, train.py:152-165);
  * the D weight-gradients of the G phase — which the reference computes and then throws
    away (train.py:159-167) — are not computed: backward(inputs=G.parameters()).
  * optionally the whole iteration is captured in a CUDA graph per (step, fading, batch) —
    with world_size > 1 the NCCL all-reduces are captured inside it, and the all-reduce of the top
    of the critic (everything above the 64 px block: most of D's parameters) is issued from a
    communication stream while the backward sweep is still working on the 64/128 px blocks.
"""
import os
import torch
import torch.distributed as dist

from . import functions as F_
from .kernels import get_kernels

_CHUNK = 2048          # elements per block of the multi-tensor Adam kernel (8 per thread)


def _d_groups(D):
    """(name, [params]) in flat-bucket order: trunk from the top of the network down, then
    the from_rgb heads, so the live set at any step is a trunk prefix + adjacent heads."""
    groups = [("linear", list(D.linear.parameters()))]
    n = D.n_layer
    for k in range(n - 1, -1, -1):
        groups.append(("progression.%d" % k, list(D.progression[k].parameters())))
    for k in range(n - 1, -1, -1):
        groups.append(("from_rgb.%d" % k, list(D.from_rgb[k].parameters())))
    return groups


def _d_active(D, step, fading):
    n = D.n_layer
    names = ["linear"] + ["progression.%d" % k for k in range(n - 1, n - 2 - step, -1)]
    names.append("from_rgb.%d" % (n - 1 - step))
    if fading and step >= 1:
        names.append("from_rgb.%d" % (n - step))
    return names


_G_RES = [8, 16, 32, 64, 128, 256]


def _g_groups(G):
    groups = [("input_layer", list(G.input_layer.parameters()) + list(G.progression_4.parameters()))]
    for r in _G_RES:
        groups.append(("progression_%d" % r, list(getattr(G, "progression_%d" % r).parameters())))
    for r in _G_RES:
        groups.append(("to_rgb_%d" % r, list(getattr(G, "to_rgb_%d" % r).parameters())))
    return groups


def _g_active(G, step, fading):
    step = min(step, G.max_step)
    names = ["input_layer"] + ["progression_%d" % r for r in _G_RES[:step]]
    names.append("to_rgb_%d" % _G_RES[step - 1])
    if fading and step >= 2:
        names.append("to_rgb_%d" % _G_RES[step - 2])
    return names


def _generic_groups(model):
    """(name, [params]) per top-level child (one per element of a ModuleList): the bucket layout
    for the model families other than train.py's Generator/Discriminator (Correct*, mnist,
    conditional).  Which groups are live for a (step, fading) is found by a probe pass."""
    groups = []
    for name, child in model.named_children():
        if isinstance(child, torch.nn.ModuleList):
            for i, sub in enumerate(child):
                ps = list(sub.parameters())
                if ps:
                    groups.append(("%s.%d" % (name, i), ps))
        else:
            ps = list(child.parameters())
            if ps:
                groups.append((name, ps))
    return groups


def _reduce_ranges(ranges, done=None, total=None):
    """The collective for a bucket's live ranges: ONE all-reduce over the span from the first live
    element to the last (the gaps hold zeros: the bucket is cleared every phase and inactive layers
    receive no gradient, so reducing them is harmless and a second NCCL launch costs more than the
    extra megabytes over NVLink), cut off below by what `done` = (a, b) already covered, with both
    ends moved outwards to 16-byte boundaries (bounded by `total` elements)."""
    rs = list(ranges)
    if done is not None:
        rs = [(max(a, done[1]), b) for a, b in rs if b > done[1]]
    if not rs:
        return []
    a, b = rs[0][0] & ~3, (rs[-1][1] + 3) & ~3
    if done is not None:
        a = max(a, done[1])
    if total is not None:
        b = min(b, total)
    return [(a, b)]


class _nvtx:
    """NVTX range around a phase of the iteration (the reference's only tracing hook is
    `profiled_function` -> record_function; here the three phases show up in nsys / ncu timelines).
    Host-side markers only: safe under CUDA-graph capture, a no-op where NVTX is unavailable."""

    def __init__(self, name):
        self.name, self.on = name, False

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
            self.on = True
        except Exception:                    # noqa: BLE001
            self.on = False

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


class FlatBucket:
    """Flat fp32 storage for the parameters, gradients and Adam state of one network."""

    def __init__(self, groups, beta1):
        self.group_params = {name: ps for name, ps in groups}
        params = [p for _, ps in groups for p in ps]
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.p = torch.empty(total, device=dev, dtype=torch.float32)
        self.g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m = torch.zeros(total, device=dev, dtype=torch.float32) if beta1 != 0.0 else None
        self.steps = torch.zeros(len(groups), device=dev, dtype=torch.float32)
        self.group_index = {}
        self.group_range = {}
        off = 0
        for gi, (name, ps) in enumerate(groups):
            start = off
            for p in ps:
                n = p.numel()
                self.p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.p[off:off + n].view(p.shape)
                p.grad = self.g[off:off + n].view(p.shape)
                off += n
            self.group_index[name] = gi
            self.group_range[name] = (start, off)
        self._plans = {}

    def plan(self, active_names):
        """Per live-set: merged contiguous ranges (for all-reduce), the chunk table and the
        step-increment mask (for multi-tensor Adam)."""
        key = tuple(active_names)
        pl = self._plans.get(key)
        if pl is None:
            dev = self.p.device
            rs = sorted(self.group_range[n] for n in active_names)
            merged = []
            for a, b in rs:
                if merged and merged[-1][1] == a:
                    merged[-1][1] = b
                else:
                    merged.append([a, b])
            chunks = []
            for n in active_names:
                a, b = self.group_range[n]
                gi = self.group_index[n]
                for s in range(a, b, _CHUNK):
                    chunks.append((s, min(_CHUNK, b - s), gi, 0))
            mask = torch.zeros(len(self.group_index), dtype=torch.float32)
            for n in active_names:
                mask[self.group_index[n]] = 1.0
            pl = dict(params=[p for n in active_names for p in self.group_params[n]],
                      ranges=[tuple(r) for r in merged],
                      chunks=torch.tensor(chunks, dtype=torch.int32, device=dev),
                      mask=mask.to(dev))
            self._plans[key] = pl
        return pl


class ProgressiveSchedule:
    """The (step, alpha) schedule of the reference loop (train.py:100-111), as an object the
    host loop calls once per iteration:

        alpha = min(1, 2 / (total_iter // n_phases) * iteration)
        when iteration > total_iter // n_phases: next resolution step (alpha, iteration = 0),
        clamped at max_step with alpha = 1.

    `next()` returns (step, alpha, new_resolution): new_resolution is True when the caller must
    rebuild its data loader at 4 * 2**step pixels (train.py:110-111).  The reference hard-codes
    3 phases and max_step = 3 (:102-109); both are parameters here."""

    def __init__(self, total_iter, init_step=1, max_step=3, n_phases=3):
        self.total_iter, self.max_step = total_iter, max_step
        self.phase_len = total_iter // n_phases
        self.step, self.iteration = init_step, 0

    def next(self):
        alpha = min(1, (2 / self.phase_len) * self.iteration)
        new_res = False
        if self.iteration > self.phase_len:
            alpha, self.iteration = 0, 0
            self.step += 1
            if self.step > self.max_step:
                alpha, self.step = 1, self.max_step
            new_res = True
        self.iteration += 1
        return self.step, alpha, new_res

    @property
    def resolution(self):
        return 4 * 2 ** self.step


class MiniStepSchedule:
    """The (step, alpha) schedule of the `proper_*` / `conditional_proper_*` loops
    (proper_cifar_train.py:77,157,165-189): every resolution step has a fade-in mini step and a
    stabilisation mini step of `images_seen_per_mini_step // batch_size` iterations each (the
    first resolution only stabilises); alpha = min(1, step_iteration / iterations_per_mini_step);
    past max_step the counter is parked (the reference's `np.inf`) and alpha stays 1.

    `next()` returns (step, alpha, new_resolution) like ProgressiveSchedule; the Correct* models
    count step 1 = 4 px, so `resolution` = 2 * 2**step."""

    def __init__(self, images_seen_per_mini_step, batch_size, init_step=1, max_step=4):
        self.per_mini = images_seen_per_mini_step // batch_size
        self.step, self.max_step = init_step, max_step
        self.step_iteration = 0

    def next(self):
        it = self.step_iteration
        alpha = 1 if it is None else min(1, it / self.per_mini)
        new_res = False
        if it is not None and ((it == self.per_mini and self.step == 1) or it == 2 * self.per_mini):
            first = self.step == 1 and it == self.per_mini
            alpha, self.step_iteration = 0, 0
            self.step += 1
            if not first and self.step > self.max_step:
                alpha, self.step_iteration, self.step = 1, None, self.max_step
            new_res = True
        if self.step_iteration is not None:
            self.step_iteration += 1
        return self.step, alpha, new_res

    @property
    def resolution(self):
        return 2 * 2 ** self.step


class Trainer:
    def __init__(self, generator, discriminator, g_running=None, lr=1e-3, betas=(0.0, 0.99),
                 eps=1e-8, ema_decay=0.999, gp_lambda=10.0, drift=0.001, process_group=None,
                 use_graph=False, segment_graphs=None, overlap_wgrad=True, overlap_passes=True,
                 n_critic=1, augment=None):
        self.G, self.D, self.G_run = generator, discriminator, g_running
        # optional ADA pipe (progan_b200.AugmentPipe) in front of every critic input - real, fake
        # and x_hat, each with its own draw, as StyleGAN2-ADA applies it; the pipe is twice
        # differentiable, so the gradient penalty differentiates through it.  Its geometric stage
        # sizes a padding from the sampled transforms on the host, so it cannot be graph-captured.
        self.augment = augment
        if augment is not None and use_graph:
            raise RuntimeError("progan_b200.Trainer: augment= needs use_graph=False (the ADA pipe's "
                               "geometric stage reads its padding margins on the host)")
        # the generator phase runs on iterations with (i + 1) % n_critic == 0 (train.py:158, 221)
        self.n_critic = int(n_critic)
        if self.n_critic < 1:
            raise ValueError("progan_b200.Trainer: n_critic must be >= 1")
        self.lr, self.betas, self.eps = lr, betas, eps
        self.ema_decay, self.gp_lambda, self.drift = ema_decay, gp_lambda, drift
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = use_graph
        # multi-GPU: ONE graph per iteration with the NCCL all-reduces captured inside it
        # (capture_error_mode="thread_local": ProcessGroupNCCL's watchdog thread queries events
        # while the capture is open, which the default global mode forbids — the round-1 attempt
        # with the default mode hung).  segment_graphs=True (or PG_SEGMENT_GRAPHS=1) keeps the
        # older form: three graphs with the collectives between the replays.
        if segment_graphs is None:
            segment_graphs = os.environ.get("PG_SEGMENT_GRAPHS", "0") == "1"
        self.segment_graphs = bool(segment_graphs) and self.world > 1
        # all-reduce of the top of the critic overlapped with the rest of its backward sweep
        self.early_reduce = os.environ.get("PG_EARLY_REDUCE", "1") == "1" and not (
            use_graph and self.segment_graphs)       # segment graphs: collectives between the replays
        self._early = None
        from .progan_modules import Discriminator as _BaseD, Generator as _BaseG
        # train.py's own models get the hand-ordered buckets (live set = two contiguous ranges);
        # every other mirrored family a generic layout + a probe pass for the live set
        self._generic = not (type(generator) is _BaseG and type(discriminator) is _BaseD)
        d_groups = _generic_groups if self._generic else _d_groups
        g_groups = _generic_groups if self._generic else _g_groups
        self._active_cache = {}
        self.bD = FlatBucket(d_groups(discriminator), betas[0])
        self.bG = FlatBucket(g_groups(generator), betas[0])
        self.bR = None
        if g_running is not None:
            self.bR = FlatBucket(g_groups(g_running), 0.0)
            for p in g_running.parameters():
                p.requires_grad_(False)
                p.grad = None
        dev = self.bD.p.device
        self.alpha_dev = torch.zeros((), device=dev, dtype=torch.float32)
        self.alpha_dev._pg_fading = True          # modules: "fade active, value lives on the device"
        self._comm_stream = torch.cuda.Stream() if (self.world > 1 and dev.type == "cuda") else None
        self.metrics = {k: torch.zeros((), device=dev, dtype=torch.float32)
                        for k in ("disc_loss", "grad_penalty", "gen_loss")}
        self.iterations = 0
        self._graphs = {}
        # weight-gradient kernels (tensor-bound) run on a side stream next to the HBM-bound kernels
        # of the data-gradient chain
        self._side = torch.cuda.Stream() if (overlap_wgrad and dev.type == "cuda") else None
        self._gp_stream = torch.cuda.Stream() if (overlap_passes and dev.type == "cuda") else None
        self._copy_stream = None
        self._ones = {}
        self._g_params = list(generator.parameters())
        self._r_params = list(g_running.parameters()) if g_running is not None else []

    # ------------------------------------------------------------------ pieces
    def _allreduce(self, bucket, plan, done=None):
        """Sum the live gradient span over the ranks (one collective, see _reduce_ranges).
        done = (a, b): that prefix of the bucket was already reduced during the backward sweep
        (_on_top_done)."""
        if self.world > 1:
            for a, b in _reduce_ranges(plan["ranges"], done, bucket.g.numel()):
                dist.all_reduce(bucket.g[a:b], group=self.pg)

    # the critic block whose OUTPUT gradient marks "every layer above is done": progression[2]
    # (64 px -> 32 px); above it sit linear + progression[3..6] = the first groups of the bucket
    _TAP_BLOCK = 2

    def _early_range(self, st):
        """(a, b) = the prefix of the D bucket that can be reduced during the backward sweep, or
        None when the feature does not apply (single GPU, generic bucket layout, low resolution
        steps whose whole critic is above the tap)."""
        if (self.world <= 1 or not self.early_reduce or self._generic or self._comm_stream is None
                or st["step"] < 4 or get_kernels().name != "cuda"):
            return None
        names = ["linear"] + ["progression.%d" % k for k in range(self.D.n_layer - 1, self._TAP_BLOCK, -1)]
        a = self.bD.group_range[names[0]][0]
        b = self.bD.group_range[names[-1]][1]
        assert a == 0 and all(self.bD.group_range[n][1] <= b for n in names)
        return (a, b)

    def _on_top_done(self, grad):
        """Tensor hook on the output of critic block _TAP_BLOCK: fires inside a backward sweep once
        the gradient of that activation is complete, i.e. after the data- and weight-gradient
        kernels of every layer above it were launched.  When both D-phase chains (gradient-penalty
        sweep, real/fake pass) have reached it, the communication stream folds those layers'
        weight-gradient workspaces into the bucket and all-reduces that prefix — while the sweeps
        go on through the 64 / 128 px blocks, where most of the time is spent."""
        e = self._early
        if e is None or not e["armed"]:
            return None
        K = get_kernels()
        cur = torch.cuda.current_stream()
        for s_ in (cur, K._side_streams.get(cur.cuda_stream)):
            if s_ is not None:
                ev = torch.cuda.Event()
                ev.record(s_)
                e["events"].append(ev)
        e["fires"] += 1
        if e["fires"] == 2:
            a, b = e["range"]
            cs = self._comm_stream
            for ev in e["events"]:
                cs.wait_event(ev)
            with torch.cuda.stream(cs):
                lo = self.bD.g.data_ptr()
                K.flush_wgrads(ptr_range=(lo + 4 * a, lo + 4 * b), join=False)
                dist.all_reduce(self.bD.g[a:b], group=self.pg)
            e["reduced"] = True
        return None

    def _adam(self, bucket, plan):
        bucket.steps.add_(plan["mask"])
        get_kernels().adam_multi(bucket.p, bucket.g, bucket.m, bucket.v, plan["chunks"], bucket.steps,
                                 self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world)
        get_kernels().refresh_packs(plan["params"])     # operand copies of the updated weights

    class _fast_paths:
        """Gradient kernels accumulate straight into the flat buckets; the conv weight gradients
        go through persistent workspaces (kernels.flush_wgrads)."""

        def __init__(self, side_stream=None):
            self.side = side_stream

        def __enter__(self):
            K = get_kernels()
            self.prev = (F_.DIRECT_GRADS, getattr(K, "defer_wgrad", False),
                         getattr(K, "wgrad_side_stream", None))
            F_.DIRECT_GRADS, K.defer_wgrad, K.wgrad_side_stream = True, True, self.side

        def __exit__(self, *exc):
            K = get_kernels()
            F_.DIRECT_GRADS, K.defer_wgrad, K.wgrad_side_stream = self.prev

    def _iteration(self, real, z, eps, step, alpha, fading, label=None, do_g=True):
        """alpha: fp32 device scalar tensor when fading else the python number.  The iteration
        is three segments separated by the two gradient all-reduces (the segments are what a
        multi-GPU run captures as CUDA graphs; the collectives stay outside the graphs)."""
        st = self._state(real, z, eps, step, alpha, fading, label, do_g)
        with self._fast_paths(self._side):
            with _nvtx("progan_b200/D phase"):
                self._seg_d(st)
                self._allreduce(self.bD, st["planD"], st.get("reduced"))
            if do_g:
                with _nvtx("progan_b200/D Adam + G phase"):
                    self._seg_g(st)
                    self._allreduce(self.bG, st["planG"])
                with _nvtx("progan_b200/G Adam + EMA"):
                    self._seg_end(st)
            else:
                with _nvtx("progan_b200/D Adam"):
                    self._seg_d_end(st)

    def _active(self, step, alpha, fading, label=None):
        """(live D groups, live G groups) for this (step, fading)."""
        if not self._generic:
            return _d_active(self.D, step, fading), _g_active(self.G, step, fading)
        key = (step, fading)
        ent = self._active_cache.get(key)
        if ent is None:
            # probe: which parameters does d D(G(z)) / d theta reach?  (batch 2, zeros; plain
            # autograd.grad with allow_unused — nothing is accumulated anywhere)
            dev = self.bD.p.device
            z = torch.zeros(2, self.G.input_dim, device=dev)
            la = (label[:1].expand(2).contiguous(),) if label is not None else ()
            prev = F_.DIRECT_GRADS
            F_.DIRECT_GRADS = False
            try:
                with torch.enable_grad():
                    out = self.D(self.G(z, *la, step=step, alpha=alpha), *la, step=step, alpha=alpha)
                    ps = list(self.D.parameters()) + list(self.G.parameters())
                    gs = torch.autograd.grad(out.sum(), ps, allow_unused=True)
            finally:
                F_.DIRECT_GRADS = prev
            used = {id(p_) for p_, g_ in zip(ps, gs) if g_ is not None}
            ent = tuple([n for n, grp in b.group_params.items() if any(id(p_) in used for p_ in grp)]
                        for b in (self.bD, self.bG))
            self._active_cache[key] = ent
        return ent

    def _state(self, real, z, eps, step, alpha, fading, label=None, do_g=True):
        namesD, namesG = self._active(step, alpha, fading, label)
        return dict(real=real, z=z, eps=eps, step=step, alpha=alpha, label=label, do_g=do_g,
                    planD=self.bD.plan(namesD), planG=self.bG.plan(namesG))

    def _aug(self, images):
        return images if self.augment is None else self.augment(images).contiguous()

    def _seg_d(self, st):
        # ---- D phase.  D(real) and D(fake) (train.py:126-139) run as ONE pass over
        # cat([real, fake]) with per-half minibatch statistics: same gradients (they accumulate
        # into the same .grad in the reference), half the launches, twice the rows per GEMM.
        K = get_kernels()
        G, D = self.G, self.D
        real, step, alpha = st["real"], st["step"], st["alpha"]
        la = (st["label"],) if st.get("label") is not None else ()      # conditional families
        la2 = (torch.cat([st["label"], st["label"]]),) if la else ()
        self.bD.g.zero_()
        B = real.shape[0]
        # critic-only iterations (n_critic > 1) never backpropagate into G: no graph is kept
        with (_NullCtx() if st.get("do_g", True) else torch.no_grad()):
            fake = G(st["z"], *la, step=step, alpha=alpha)
        # The gradient-penalty pass (train.py:142-151) and the real/fake pass are independent
        # given `fake`: they run as two concurrent chains (streams; parallel branches of the
        # captured graph).  Every gradient kernel accumulates atomically (workspaces, bias sums),
        # so the two chains may add into the same buckets at the same time; the latency-bound
        # low-resolution layers of one chain hide behind the large kernels of the other.
        s2 = self._gp_stream if (self._gp_stream is not None and K.name == "cuda" and K.conv_impl == "tc") else None
        cur = torch.cuda.current_stream() if s2 is not None else None
        early = self._early_range(st)
        if early is not None:
            self._early = dict(armed=False, fires=0, events=[], range=early, reduced=False)
            D._bwd_tap = (self._TAP_BLOCK, self._on_top_done)
        if s2 is not None:
            s2.wait_stream(cur)
        with (torch.cuda.stream(s2) if s2 is not None else _NullCtx()):
            x_hat = K.interp_xhat(real, fake.detach(), st["eps"].reshape(-1)).requires_grad_(True)
            hat = D(self._aug(x_hat), *la, step=step, alpha=alpha)
            # d(sum_n hat_n)/dx_hat (train.py:146) with an explicit unit seed: no sum kernel and no
            # expand in its backward
            ones = self._ones.get(tuple(hat.shape))
            if ones is None:
                ones = self._ones[tuple(hat.shape)] = torch.ones_like(hat)
            (g,) = torch.autograd.grad(outputs=hat, inputs=x_hat, grad_outputs=ones, create_graph=True)
            gp = F_.gradient_penalty(g, self.gp_lambda)
            if early is not None:
                self._early["armed"] = True          # the first-order sweep above must not count
            gp.backward()
        both = D(self._aug(torch.cat([real, fake.detach()])), *la2, step=step, alpha=alpha, mbstd_group=B)
        # loss value (-> metric) and dL/d(outputs) in one launch, then backward from that seed
        both.backward(K.wgan_loss(both.detach(), B, self.drift, self.metrics["disc_loss"]))
        if s2 is not None:
            cur.wait_stream(s2)
        if early is not None:
            D._bwd_tap = None
            if self._early["reduced"]:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
                st["reduced"] = early
            self._early = None
        K.flush_wgrads()
        st["fake"] = fake
        st["gp"] = gp.detach()

    def _seg_d_end(self, st):
        # critic-only iteration: D update and the metric, no generator phase (train.py:155-158)
        self._adam(self.bD, st["planD"])
        self.metrics["grad_penalty"].add_(st["gp"])

    def _seg_g(self, st):
        K = get_kernels()
        self._adam(self.bD, st["planD"])
        self.metrics["grad_penalty"].add_(st["gp"])
        # ---- G phase (D already updated, train.py:158-169)
        self.bG.g.zero_()
        la = (st["label"],) if st.get("label") is not None else ()
        out = self.D(self._aug(st["fake"]), *la, step=st["step"], alpha=st["alpha"])
        out.backward(K.wgan_loss(out.detach(), 0, 0.0, self.metrics["gen_loss"]), inputs=st["planG"]["params"])
        K.flush_wgrads()

    def _seg_end(self, st):
        K = get_kernels()
        self._adam(self.bG, st["planG"])
        if self.bR is not None:
            K.ema(self.bR.p, self.bG.p, self.ema_decay)
            K.drop_packs(self._r_params)

    # ------------------------------------------------------------------ public
    def _stage_host_inputs(self, real, z, eps):
        """Host (pinned) inputs -> device, on a copy stream and into one of two staging sets, so
        the H2D transfer of iteration i+1 overlaps the kernels of iteration i (the host runs
        ahead of the GPU); the compute stream only waits for the copy's event."""
        dev = self.bD.p.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging, self._stg_free, self._stg_idx = {}, {}, 0
        self._stg_idx ^= 1
        key = (self._stg_idx, tuple(real.shape), tuple(z.shape))
        stg = self._staging.get(key)
        if stg is None:
            stg = tuple(torch.empty(t.shape, device=dev, dtype=t.dtype) for t in (real, z, eps))
            self._staging[key] = stg
        cs = self._copy_stream
        free = self._stg_free.get(key)
        if free is not None:
            cs.wait_event(free)              # the compute stream has consumed this set
        with torch.cuda.stream(cs):
            for d, h in zip(stg, (real, z, eps)):
                d.copy_(h, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cs)
        cur = torch.cuda.current_stream()
        cur.wait_event(done)
        self._stg_key = key
        return stg

    def step(self, real, z, eps, step, alpha, label=None):
        """One full iteration: real [B,3,R,R] fp32, z [B,zdim], eps [B,1,1,1] (drawn by the caller
        on the CPU generator, train.py:133,142) — device tensors, or host tensors (pinned for an
        asynchronous copy), which are then transferred on a copy stream."""
        host = (not real.is_cuda) and self.bD.p.is_cuda
        if host:
            real, z, eps = self._stage_host_inputs(real, z, eps)
        if label is not None and not label.is_cuda and self.bD.p.is_cuda:
            label = label.to(self.bD.p.device, non_blocking=True)
        self._step_device(real, z, eps, step, alpha, label)
        if host:                             # staging set free again once this iteration has read it
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stg_free[self._stg_key] = ev

    def _step_device(self, real, z, eps, step, alpha, label=None):
        fading = 0 <= alpha < 1
        if fading:
            self.alpha_dev.fill_(float(alpha))
        a = self.alpha_dev if fading else alpha
        do_g = (self.iterations + 1) % self.n_critic == 0
        if not self.use_graph:
            self._iteration(real, z, eps, step, a, fading, label, do_g)
        else:
            self._graph_step(real, z, eps, step, a, fading, label, do_g)
        self.iterations += 1

    def _graph_step(self, real, z, eps, step, a, fading, label=None, do_g=True):
        key = (step, fading, tuple(real.shape), label is not None, do_g)
        ent = self._graphs.get(key)
        K = get_kernels()
        if ent is None:
            sreal, sz, seps = real.clone(), z.clone(), eps.clone()
            slabel = label.clone() if label is not None else None
            # warm up on a side stream (allocator + lazy init), restoring state afterwards
            state = [self.bD.p, self.bD.v, self.bD.steps, self.bG.p, self.bG.v, self.bG.steps]
            state += [b.m for b in (self.bD, self.bG) if b.m is not None]
            state += list(self.metrics.values())
            snap = [b.clone() for b in state]
            snapR = self.bR.p.clone() if self.bR is not None else None
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._iteration(sreal, sz, seps, step, a, fading, slabel, do_g)
            torch.cuda.current_stream().wait_stream(s)
            for dst, src in zip(state, snap):
                dst.copy_(src)
            if snapR is not None:
                self.bR.p.copy_(snapR)
            # operand copies follow the restored weights (same storage: the capture below and
            # every replay keep them consistent through the refresh launches after each Adam)
            K.refresh_packs(list(self.D.parameters()))
            K.refresh_packs(list(self.G.parameters()))
            torch.cuda.synchronize()
            if not self.segment_graphs:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    self._iteration(sreal, sz, seps, step, a, fading, slabel, do_g)
                graphs = [graph]
            else:
                # one graph per segment; the NCCL all-reduces run between the replays
                st = self._state(sreal, sz, seps, step, a, fading, slabel, do_g)
                graphs, pool = [], None
                for seg in ((self._seg_d, self._seg_g, self._seg_end) if do_g
                            else (self._seg_d, self._seg_d_end)):
                    gph = torch.cuda.CUDAGraph()
                    with self._fast_paths(self._side), torch.cuda.graph(gph, pool=pool,
                                                              capture_error_mode="thread_local"):
                        seg(st)
                    pool = gph.pool()
                    graphs.append(gph)
                del st
            namesD, namesG = self._active(step, a, fading, slabel)
            ent = (graphs, sreal, sz, seps, self.bD.plan(namesD), self.bG.plan(namesG), slabel)
            self._graphs[key] = ent
            # the capture itself did not execute anything
        graphs, sreal, sz, seps, planD, planG, slabel = ent
        if slabel is not None:
            slabel.copy_(label, non_blocking=True)
        sreal.copy_(real, non_blocking=True)
        sz.copy_(z, non_blocking=True)
        seps.copy_(eps, non_blocking=True)
        if len(graphs) == 1:
            graphs[0].replay()
        else:
            graphs[0].replay()
            self._allreduce(self.bD, planD)
            graphs[1].replay()
            if do_g:
                self._allreduce(self.bG, planG)
                graphs[2].replay()
        if do_g and self._r_params:
            # the captured EMA kernel rewrote g_running's weights behind autograd's back: forget
            # their operand copies so the next g_running(...) call packs the current values
            K.drop_packs(self._r_params)

    def read_metrics(self, reset=True):
        """One host sync for all three running sums (the reference syncs three times per
        iteration with .item(), train.py:152,153,165)."""
        vals = torch.stack(list(self.metrics.values())).tolist()
        out = dict(zip(self.metrics.keys(), vals))
        if reset:
            for m in self.metrics.values():
                m.zero_()
        return out
