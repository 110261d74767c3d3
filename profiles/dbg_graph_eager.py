import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, common, helpers, progan_b200
from test_gpu_train import _run
for rep in range(3):
    for name in ["s2_a0.5", "s3_a0.25", "s5_a0.5"]:
        for iters in (1, 3):
            _, e1, *_ = _run(name, "bf16", "tc", False, iters)
            _, e2, *_ = _run(name, "bf16", "tc", False, iters)
            _, g1, *_ = _run(name, "bf16", "tc", True, iters)
            m1, m2, m3 = e1.read_metrics(), e2.read_metrics(), g1.read_metrics()
            f = lambda m: " ".join("%.4f" % v for v in m.values())
            print(rep, name, iters, "| e-e D %.1e g-e D %.1e |" % (helpers.rel(e2.bD.p, e1.bD.p), helpers.rel(g1.bD.p, e1.bD.p)), f(m1), "|", f(m2), "|", f(m3))
