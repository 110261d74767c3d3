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
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


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
    # launches of an iteration and times the dominant kernel live.  Every 3x3 tensor-core conv
    # launch (conv4_tc_kernel: forward, data-gradient and GP tangent convs) is issued three
    # times back to back between two CUDA events on its stream — the call is idempotent, and the
    # repeats hide the host launch latency of eager mode — and a third of the time is charged.
    conv_recs = []
    orig_call = K._call

    def timed_conv_call(name, *a):
        if name == "pg_conv_tc" and a[11] == 9 and a[6] % 16 == 0 and a[7] % 8 == 0 and a[9] == a[10] \
                and a[10] in (32, 64, 128):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            orig_call(name, *a)                      # the launch that belongs to the iteration
            e0.record()
            for _ in range(3):
                orig_call(name, *a)
            e1.record()
            K.launches -= 3
            conv_recs.append((2.0 * a[5] * a[6] * a[7] * a[8] * a[9] * 9, e0, e1))
        else:
            orig_call(name, *a)

    K2 = K.launches
    was_graph = tr.use_graph
    tr.use_graph = False
    K._call = timed_conv_call
    try:
        tr.step(real, z, eps, step, alpha)
        torch.cuda.synchronize()
    finally:
        K._call = orig_call
        tr.use_graph = was_graph
    eager_launches = K.launches - K2
    conv_flops = sum(f for f, _, _ in conv_recs)
    conv_ms = sum(e0.elapsed_time(e1) / 3.0 for _, e0, e1 in conv_recs)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = B * world * args.steps
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    peak_tf, peak_hbm, which = peaks()
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
                   "l2": "working set %.0f MB per pass > 126 MB L2" % (B * res * res * 64 * 2 * 4 / 1e6)},
        "e2e": {"value": round(e2e, 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 12},
        "gpu_launches": eager_launches * args.steps,
        "roofline": roofline(conv_flops, conv_ms, len(conv_recs), ms / args.steps, achieved_tf, res),
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(res, B, alpha, sample_batch=args.cpu_batch)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline(conv_flops, conv_ms, n_conv, step_ms, step_tf, res):
    """Roofline of the dominant kernel (conv4_tc_kernel, tensor-bound): algorithmic FLOPs of its
    launches in one iteration over their summed device time, against the measured sustained bf16
    peak (the kernel runs inside a long step).  `traffic`: DRAM bytes of the largest launch
    (128->128 @64px, batch 128) from the ncu --set full capture in profiles/ (algorithmic
    268 MB: no wasted re-reads; part of the output is still in L2 when the kernel ends)."""
    peak_tf, _, which = peaks()
    ach = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "conv4_tc_kernel", "achieved": round(ach, 2), "peak": peak_tf,
            "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
            "traffic": 223.7e6 if res == 128 else None,
            "launches_per_step": n_conv, "kernel_ms_per_step": round(conv_ms, 3),
            "kernel_share_of_step": round(conv_ms / step_ms, 3) if step_ms > 0 else None,
            "whole_step": {"achieved": round(step_tf, 2), "frac": round(step_tf / peak_tf, 4),
                           "note": "14 F_D + 3 F_G = %.2f GFLOP/img over the step time"
                                   % (step_flops(res) / 1e9)},
            "note": "algorithmic conv FLOPs of the kernel's launches / their device time (CUDA events); "
                    "peak = %s; ncu tensor-pipe active 65-71%% (profiles/)" % which}


def cpu_baseline(res, B, alpha, sample_batch=32, iters=2):
    """The oracle (fp32 PyTorch restatement of the reference loop, pinned to the reference by
    the golden vectors) timed on the host cores, on a bounded sample of the workload."""
    from oracle import progan_oracle as O
    import progan_b200
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[res]
    torch.manual_seed(0)
    with torch.device("cpu"):
        G = progan_b200.Generator(128, 128, tanh=False)
        D = progan_b200.Discriminator(128)
    PG, PD, PR = O.params_of(G), O.params_of(D), O.params_of(G, False)
    optG, optD = O.AdamState(PG), O.AdamState(PD)
    real, z, eps = make_inputs(sample_batch, res, 128, 1234)
    O.train_iteration(PG, PD, PR, optG, optD, real, z, eps, step, alpha)      # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        O.train_iteration(PG, PD, PR, optG, optD, real, z, eps, step, alpha)
    dt = (time.perf_counter() - t0) / iters
    return {"value": round(sample_batch / dt, 3), "unit": "img/s", "cores": threads, "kind": "port",
            "sample": "%d full iterations at batch %d, %dpx (1 warm-up), oracle/progan_oracle.py"
                      % (iters, sample_batch, res)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    res, B = args.res, args.batch
    step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[res]
    cb = cpu_baseline(res, B, args.alpha, sample_batch=args.cpu_batch, iters=max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": "G+D+GP train imgs/sec at %dpx" % res,
            "value": cb["value"], "unit": "img/s", "n_gpus": world, "steps": max(1, min(args.steps, 3)),
            "warmup": 1, "ms_per_step": round(1e3 * args.cpu_batch / cb["value"], 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": "train.py CelebA-shape G(128,128)/D(128) step %d (%dpx) alpha=%.2f, "
                                   "bounded sample: batch %d per step on the host cores"
                                   % (step, res, args.alpha, args.cpu_batch)},
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
        run_product(a)
