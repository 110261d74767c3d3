"""Data-parallel host logic on CPU: world_size 2 over gloo, kernels emulated.
Each rank runs Trainer.step on its own shard; the all-reduced, averaged gradients must make
both replicas identical and equal to Adam applied to the mean of the two ranks' gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, name):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import common
    import helpers
    import progan_b200
    from emul_kernels import EmulKernels
    progan_b200.set_kernels(EmulKernels())
    inp = common.make_inputs(name)
    g = torch.Generator().manual_seed(1234 + rank)          # per-rank shard
    real = torch.rand(inp["real"].shape, generator=g) * 2 - 1
    z = torch.randn(inp["z"].shape, generator=g)
    eps = torch.rand(inp["eps"].shape, generator=g)
    label = torch.randint(0, inp["num_classes"], inp["label"].shape, generator=g) if inp["num_classes"] else None
    G, D = helpers.build_models(inp, "fp32", name=name)
    Gr, _ = helpers.build_models(inp, "fp32", name=name)
    tr = progan_b200.Trainer(G, D, Gr)
    assert tr.world == world
    tr.step(real, z, eps, inp["step"], inp["alpha"], label=label)
    torch.save(dict(pD=tr.bD.p.clone(), pG=tr.bG.p.clone(), gD=tr.bD.g.clone(), gG=tr.bG.g.clone(),
                    pR=tr.bR.p.clone(), real=real, z=z, eps=eps, label=label),
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["s2_a0.5", "k2_a0.5_eq", "p2_a0.5"])
def test_two_rank_data_parallel_step(tmp_path, name):
    """train.py's models (hand-ordered buckets), a class-conditional family (label plane, generic
    buckets) and a projection critic, each rank with its own images, latents and labels."""
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), name), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # replicas identical after the step
    for k in ("pD", "pG", "pR", "gD", "gG"):
        assert torch.equal(r0[k], r1[k]), k
    # and equal to a single-process computation on the mean gradient
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import common
    import helpers
    import progan_b200
    from emul_kernels import EmulKernels
    prev = progan_b200.set_kernels(EmulKernels())
    try:
        inp = common.make_inputs(name)
        gsum = None
        for r in (r0, r1):
            G, D = helpers.build_models(inp, "fp32", name=name)
            res, _ = helpers.product_train_step(G, D, r["real"], r["z"], r["eps"], inp["step"], inp["alpha"],
                                                label=r["label"])
            tr = progan_b200.Trainer(G, D)          # flat view of the D grads in bucket order
            flat = torch.zeros_like(tr.bD.g)
            for gname, (a, b) in tr.bD.group_range.items():
                off = a
                for p in tr.bD.group_params[gname]:
                    key = [k for k, q in D.named_parameters() if q is p][0]
                    if key in res["d_grads"]:
                        flat[off:off + p.numel()] = res["d_grads"][key].reshape(-1)
                    off += p.numel()
            gsum = flat if gsum is None else gsum + flat
        assert helpers.rel(r0["gD"], gsum) < 1e-5      # all-reduce(SUM); Adam divides by world
    finally:
        progan_b200.set_kernels(prev)
