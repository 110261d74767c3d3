"""Run-to-run reproducibility of the gradient buckets (same weights, same inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, common, helpers, progan_b200
DEV = "cuda"
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True


def grads(name, batched=True):
    inp = common.make_inputs(name)
    G, D = helpers.build_models(inp, "bf16", device=DEV)
    tr = progan_b200.Trainer(G, D, None)
    snaps = []
    orig = tr._adam
    def spy(bucket, plan):
        snaps.append(bucket.g.clone())
        orig(bucket, plan)
    tr._adam = spy
    real, z, eps = inp["real"].to(DEV), inp["z"].to(DEV), inp["eps"].to(DEV)
    tr.step(real, z, eps, inp["step"], inp["alpha"])
    torch.cuda.synchronize()
    return tr, snaps


for name in ["s2_a0.5", "s3_a0.25", "s5_a0.5"]:
    for rep in range(3):
        t1, s1 = grads(name)
        t2, s2 = grads(name)
        print(name, rep, "gD rel %.2e  gG rel %.2e" % (helpers.rel(s2[0], s1[0]), helpers.rel(s2[1], s1[1])))
        if helpers.rel(s2[0], s1[0]) > 1e-4:
            for gname, (a, b) in t1.bD.group_range.items():
                if s1[0][a:b].abs().max() > 0:
                    print("    D", gname, "rel %.2e" % helpers.rel(s2[0][a:b], s1[0][a:b]))
        if helpers.rel(s2[1], s1[1]) > 1e-4:
            for gname, (a, b) in t1.bG.group_range.items():
                if s1[1][a:b].abs().max() > 0:
                    print("    G", gname, "rel %.2e" % helpers.rel(s2[1][a:b], s1[1][a:b]))
