import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, common, helpers, progan_b200
from test_gpu_train import _run
for name in ["s2_a0.5", "s5_a0.5"]:
    for iters in (1, 2, 3):
        _, e, *_ = _run(name, "bf16", "tc", False, iters)
        _, g, *_ = _run(name, "bf16", "tc", True, iters)
        _, s, *_ = _run(name, "bf16", "tc", True, iters, True)
        print(name, iters, "graph-eager D %.1e G %.1e | seg-eager D %.1e G %.1e R %.1e" % (
            helpers.rel(g.bD.p, e.bD.p), helpers.rel(g.bG.p, e.bG.p), helpers.rel(s.bD.p, e.bD.p), helpers.rel(s.bG.p, e.bG.p), helpers.rel(s.bR.p, e.bR.p)),
            e.read_metrics(), s.read_metrics())
        if helpers.rel(s.bG.p, e.bG.p) > 1e-5:
            for gname, (a, b) in e.bG.group_range.items():
                print("    ", gname, "%.2e" % helpers.rel(s.bG.p[a:b], e.bG.p[a:b]), "steps", float(s.bG.steps[s.bG.group_index[gname]]), float(e.bG.steps[e.bG.group_index[gname]]))
