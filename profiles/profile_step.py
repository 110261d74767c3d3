"""One eager training iteration between cudaProfilerStart/Stop, for ncu:
  python profiles/profile_step.py [--res 128] [--batch 64]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import progan_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--alpha", type=float, default=0.5)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5, 256: 6}[a.res]
K = progan_b200.get_kernels()
K.conv_impl, K.wgrad_tc = "tc", True
torch.manual_seed(0)
G = progan_b200.Generator(128, 128, tanh=False).to(dev)
D = progan_b200.Discriminator(128).to(dev)
Gr = progan_b200.Generator(128, 128, tanh=False).to(dev)
tr = progan_b200.Trainer(G, D, Gr, use_graph=False)
g = torch.Generator().manual_seed(1234)
real = (torch.rand(a.batch, 3, a.res, a.res, generator=g) * 2 - 1).to(dev)
z = torch.randn(a.batch, 128, generator=g).to(dev)
eps = torch.rand(a.batch, 1, 1, 1, generator=g).to(dev)
for _ in range(a.warmup):
    tr.step(real, z, eps, step, a.alpha)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(real, z, eps, step, a.alpha)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", tr.read_metrics())
