import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, progan_b200
dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = progan_b200.Generator(128, 128, tanh=False).to(dev); D = progan_b200.Discriminator(128).to(dev)
Gr = progan_b200.Generator(128, 128, tanh=False).to(dev)
tr = progan_b200.Trainer(G, D, Gr, use_graph=True)
g = torch.Generator().manual_seed(1)
real_h = (torch.rand(64, 3, 128, 128, generator=g) * 2 - 1).pin_memory(); z_h = torch.randn(64, 128, generator=g).pin_memory(); eps_h = torch.rand(64, 1, 1, 1, generator=g).pin_memory()
real, z, eps = real_h.to(dev), z_h.to(dev), eps_h.to(dev)
loss_h = torch.zeros(3).pin_memory()
def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3
def a(): tr.step(real, z, eps, 5, 0.5)
def b(): tr.step(real_h, z_h, eps_h, 5, 0.5)
def c():
    tr.step(real_h, z_h, eps_h, 5, 0.5); loss_h.copy_(torch.stack(list(tr.metrics.values())), non_blocking=True)
def d():
    r = real_h.to(dev, non_blocking=True); tr.step(r, z, eps, 5, 0.5)
for name, fn in (("device inputs", a), ("host inputs (copy stream)", b), ("host inputs + loss readback", c), ("h2d on main stream", d), ("device inputs again", a)):
    gpu, cpu = timed(fn)
    print("%-32s gpu %.3f ms/step   cpu enqueue %.3f ms/step" % (name, gpu, cpu))
