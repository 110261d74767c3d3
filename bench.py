"""bench.py — G+D+GP train imgs/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res 128] [--batch 64]
  python bench.py --impl reference ...     # the reference's CPU path on the host cores

One "step" = one full iteration of the reference hot loop (train.py:97-169, n_critic=1):
D(real), G(z), D(fake), gradient penalty with double backward, D Adam, D(G(z)) -> G
backward, G Adam, EMA.  Workload = BASELINE config 4: Generator(128,128,tanh=False) /
Discriminator(128), batch 64 per GPU, 128 px (step 5), alpha = 0.5 (fade active, both heads
run).  Synthetic images U(-1,1), random-init weights.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# forward conv FLOPs per image (probed on the reference modules), channel=128
F_G = {8: 0.0478e9, 16: 0.199e9, 32: 0.804e9, 64: 1.711e9, 128: 2.619e9, 256: 5.046e9}
F_D = {8: 0.0431e9, 16: 0.194e9, 32: 0.799e9, 64: 2.612e9, 128: 4.427e9, 256: 6.854e9}


def step_flops(res):
    """algorithmic conv FLOPs per image per train step = 14 F_D + 3 F_G."""
    return 14 * F_D[res] + 3 * F_G[res]


def peaks():
    """(burst bf16 TF/s, sustained bf16 TF/s, HBM GB/s, provenance)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops"], p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def file_sha16(path):
    import hashlib
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full summary
    (profiles/r2/conv4_ncu_summary.json) — only when that capture was taken from the kernel
    source this process runs (sha of csrc/conv4_tc.cu), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2", "conv4_ncu_summary.json")) as f:
            rec = json.load(f)
        src = os.path.join(ROOT, "progressive-gan-pytorch_b200", "csrc", "conv4_tc.cu")
        if rec.get("conv4_tc_cu_sha16") != file_sha16(src):
            return None, "profiles/r2/conv4_ncu_summary.json is from another build of conv4_tc.cu"
        return rec["dram_bytes_per_launch"], rec.get("note", "")
    except Exception as e:                                     # noqa: BLE001
        return None, "no ncu summary (%s)" % type(e).__name__


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def arm_deadline(seconds=None):
    """Hang guard for the product arm: if the process is still alive after `seconds`
    (PG_BENCH_DEADLINE_S, default 1500; 0 disables) every thread's Python stack goes to stderr and
    the process exits with status 1 — under torchrun that takes the other ranks down with it —
    instead of holding the GPUs until an outer timeout kills the box."""
    import faulthandler
    if seconds is None:
        seconds = float(os.environ.get("PG_BENCH_DEADLINE_S", "1500"))
    if seconds > 0:
        faulthandler.dump_traceback_later(seconds, exit=True)
    return seconds


def make_inputs(batch, res, zdim, seed):
    g = torch.Generator().manual_seed(seed)                      # CPU generator (train.py:133,142)
    real = (torch.rand(batch, 3, res, res, generator=g) * 2 - 1)
    z = torch.randn(batch, zdim, generator=g)
    eps = torch.rand(batch, 1, 1, 1, generator=g)
    return real, z, eps


def run_product(args):
    import torch.distributed as dist
    import progan_b200
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res, B = args.res, args.batch
    step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[res]
    alpha = args.alpha
    K = progan_b200.get_kernels()
    K.conv_impl, K.wgrad_tc = args.conv, args.conv == "tc"
    torch.manual_seed(0)
    G = progan_b200.Generator(128, 128, pixel_norm=True, tanh=False, precision=args.precision).to(dev)
    D = progan_b200.Discriminator(128, precision=args.precision).to(dev)
    Grun = progan_b200.Generator(128, 128, pixel_norm=True, tanh=False, precision=args.precision).to(dev)
    Grun.load_state_dict(G.state_dict())
    if world > 1:
        for p in list(G.parameters()) + list(D.parameters()) + list(Grun.parameters()):
            dist.broadcast(p.data, 0)
    tr = progan_b200.Trainer(G, D, Grun, use_graph=not args.no_graph)
    real_h, z_h, eps_h = [t.pin_memory() for t in make_inputs(B, res, 128, 1234 + rank)]
    real, z, eps = real_h.to(dev), z_h.to(dev), eps_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def dev_step():
        tr.step(real, z, eps, step, alpha)

    loss_h = torch.zeros(3).pin_memory()

    def e2e_step():
        # public API with HOST buffers: H2D of this step's inputs (pinned memory; Trainer.step
        # issues the copies on its copy stream) + D2H of the step's losses, every step
        tr.step(real_h, z_h, eps_h, step, alpha)
        loss_h.copy_(torch.stack(list(tr.metrics.values())), non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.launches
    ms = timed(dev_step, args.steps)
    launches = K.launches - l0
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True
    # One more EAGER iteration on every rank (it contains the gradient all-reduces): counts the
    # launches of an iteration and records every launch served by the dominant kernel
    # (conv4_tc_kernel: 3x3 forward, data-gradient and GP tangent convs, with and without the fused
    # activation-backward epilogue).  Afterwards, with nothing else in flight, each recorded launch
    # is re-issued three times back to back between two CUDA events on its stream (same operands,
    # same kernel variant; the call is idempotent) and a third of the time is charged: the kernel
    # timed ALONE, which is what the measured burst peak is the denominator for.
    conv_calls = []
    orig_call = K._call

    def recording_call(name, *a):
        dims = None         # (N, H, W, Cin, Cout) of a launch served by conv4_tc_kernel
        if name == "pg_conv_tc" and a[11] == 9 and a[6] % 16 == 0 and a[7] % 8 == 0 and a[9] == a[10] \
                and a[10] in (32, 64, 128):
            dims = a[5:10]
            conv_calls.append((name, a, dims))
        elif name == "pg_conv_tc_actbwd":            # data-gradient conv with the fused act-backward
            dims = a[3:8]
            conv_calls.append((name, a[:13] + (None,) + a[14:], dims))   # repeats: no bias-gradient add
        orig_call(name, *a)

    K2 = K.launches
    was_graph = tr.use_graph
    tr.use_graph = False
    K._call = recording_call
    try:
        tr.step(real, z, eps, step, alpha)
        torch.cuda.synchronize()
    finally:
        K._call = orig_call
        tr.use_graph = was_graph
    eager_launches = K.launches - K2
    conv_recs = []
    st_now = torch.cuda.current_stream().cuda_stream
    # the host needs ~25 us per call and many of these launches run 10-30 us: a spin kernel keeps
    # the GPU busy while the whole sequence is enqueued, so no timed launch waits for the host
    torch.cuda._sleep(int(60e6))
    for name, a, dims in conv_calls:
        a = a[:-1] + (st_now,)                       # everything on the current stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        orig_call(name, *a)                          # warm (descriptor cache, L2 state of a chain)
        e0.record()
        for _ in range(3):
            orig_call(name, *a)
        e1.record()
        n_, h_, w_, ci_, co_ = dims
        conv_recs.append((2.0 * n_ * h_ * w_ * ci_ * co_ * 9, e0, e1))
    torch.cuda.synchronize()
    K.launches = K2 + eager_launches
    conv_flops = sum(f for f, _, _ in conv_recs)
    conv_ms = sum(e0.elapsed_time(e1) / 3.0 for _, e0, e1 in conv_recs)
    # data-parallel correctness on the real NCCL path: after the timed steps every rank must hold
    # bit-identical parameters (same all-reduced gradients -> same Adam/EMA updates)
    replicas = None
    if world > 1:
        sums = torch.stack([tr.bD.p.double().sum(), tr.bG.p.double().sum(), tr.bR.p.double().sum(),
                            tr.bD.p.double().abs().sum(), tr.bG.p.double().abs().sum()])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        replicas = all(torch.equal(allsums[0], t) for t in allsums)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = B * world * args.steps
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    achieved_tf = (value / world) * step_flops(res) / 1e12
    h2d = real_h.numel() * 4 + z_h.numel() * 4 + eps_h.numel() * 4
    line = {
        "metric": "G+D+GP train imgs/sec at %dpx" % res, "value": round(value, 2), "unit": "img/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "train.py CelebA-shape G(128,128)/D(128) step %d (%dpx) alpha=%.2f "
                               "batch %d/GPU, full iteration incl. Adam+EMA" % (step, res, alpha, B),
                   "parallelism": "dp%d" % world, "conv": args.conv, "cuda_graph": not args.no_graph,
                   "allreduce": (("captured in the iteration graph" +
                                  ("; top of the critic reduced during the backward sweep"
                                   if tr.early_reduce and step >= 4 else ""))
                                 if world > 1 and not tr.segment_graphs
                                 else ("between segment graphs" if world > 1 else None)),
                   "l2": "working set %.0f MB per pass > 126 MB L2" % (B * res * res * 64 * 2 * 4 / 1e6)},
        "e2e": {"value": round(e2e, 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 12},
        "gpu_launches": eager_launches * args.steps,
        "roofline": roofline(conv_flops, conv_ms, len(conv_recs), ms / args.steps, achieved_tf, res,
                             ms / 1e3),
        "clocks": sampler.summary(),
    }
    if replicas is not None:
        line["replicas_identical"] = bool(replicas)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(res, B, alpha, sample_batch=args.cpu_batch)[0]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline(conv_flops, conv_ms, n_conv, step_ms, step_tf, res, timed_s):
    """Roofline of the dominant kernel (conv4_tc_kernel, tensor-bound): algorithmic FLOPs of its
    launches in one iteration over their summed device time (CUDA events on the launching
    stream).  `peak` is the measured BURST bf16 figure when the timed region is shorter than a
    second (the kernel is effectively timed alone, clocks at maximum), the sustained one for a long
    region; both fractions are reported.  `traffic`: DRAM bytes of one launch from the committed
    ncu --set full capture of THIS kernel source (None when the capture is from another build)."""
    burst, sustained, _, which = peaks()
    use_burst = timed_s < 1.0
    peak_tf = burst if use_burst else sustained
    ach = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, tnote = ncu_traffic() if res == 128 else (None, "captured at 128 px only")
    return {"bound": "tensor", "kernel": "conv4_tc_kernel", "achieved": round(ach, 2), "peak": peak_tf,
            "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
            "peak_kind": "burst" if use_burst else "sustained",
            "frac_of_burst": round(ach / burst, 4), "frac_of_sustained": round(ach / sustained, 4),
            "traffic": traffic, "traffic_note": tnote,
            "launches_per_step": n_conv, "kernel_ms_per_step": round(conv_ms, 3),
            "kernel_share_of_step": round(conv_ms / step_ms, 3) if step_ms > 0 else None,
            "whole_step": {"achieved": round(step_tf, 2), "frac": round(step_tf / peak_tf, 4),
                           "frac_of_burst": round(step_tf / burst, 4),
                           "frac_of_sustained": round(step_tf / sustained, 4),
                           "note": "14 F_D + 3 F_G = %.2f GFLOP/img over the step time"
                                   % (step_flops(res) / 1e9)},
            "note": "algorithmic conv FLOPs of the kernel's launches of one iteration / their device time "
                    "(CUDA events; every launch of the iteration re-issued alone, 3x back to back); "
                    "peaks %s" % which}


def cpu_baseline(res, B, alpha, sample_batch=32, iters=2):
    """The reference's CPU path timed on the host cores, on a bounded sample of the workload: the
    UNMODIFIED reference modules staged under baseline/_ref/ (oracle/stage_reference.py) driven by
    the loop body of train.py:97-169 (oracle/ref_loop.py) — kind "reference"; when they are not
    staged, the oracle restatement (oracle/progan_oracle.py, pinned to the reference's golden
    vectors) — kind "port"."""
    from oracle import ref_loop
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[res]
    real, z, eps = make_inputs(sample_batch, res, 128, 1234)
    R, where = ref_loop.import_reference()
    torch.manual_seed(0)
    if R is not None:
        G, D, Grun, g_opt, d_opt = ref_loop.build(R, 128, 128)

        def it():
            ref_loop.iteration(G, D, Grun, g_opt, d_opt, real, z, eps, step, alpha)
        kind, what = "reference", "unmodified progan_modules.py from %s + train.py:97-169 loop body" % (
            "baseline/_ref" if "baseline" in where else where)
    else:
        from oracle import progan_oracle as O
        import progan_b200
        with torch.device("cpu"):
            G = progan_b200.Generator(128, 128, tanh=False)
            D = progan_b200.Discriminator(128)
        PG, PD, PR = O.params_of(G), O.params_of(D), O.params_of(G, False)
        optG, optD = O.AdamState(PG), O.AdamState(PD)

        def it():
            O.train_iteration(PG, PD, PR, optG, optD, real, z, eps, step, alpha)
        kind, what = "port", "oracle/progan_oracle.py"
    it()                                                                     # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        it()
    dt = (time.perf_counter() - t0) / iters
    return {"value": round(sample_batch / dt, 3), "unit": "img/s", "cores": threads, "kind": kind,
            "sample": "%d full iterations at batch %d, %dpx (1 warm-up), fp32, %s"
                      % (iters, sample_batch, res, what)}, dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores,
    at the product arm's config (batch 64 per step by default)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    res, B = args.res, args.batch
    step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[res]
    iters = max(1, min(args.steps, 2))
    cb, dt = cpu_baseline(res, B, args.alpha, sample_batch=B, iters=iters)
    line = {"impl": "reference", "metric": "G+D+GP train imgs/sec at %dpx" % res,
            "value": cb["value"], "unit": "img/s", "n_gpus": world, "steps": iters,
            "warmup": 1, "ms_per_step": round(1e3 * dt, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": "train.py CelebA-shape G(128,128)/D(128) step %d (%dpx) alpha=%.2f "
                                   "batch %d, full iteration incl. Adam+EMA, on the host cores"
                                   % (step, res, args.alpha, B)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--conv", default="tc", choices=["tc", "simt"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=32)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        arm_deadline()
        run_product(a)
