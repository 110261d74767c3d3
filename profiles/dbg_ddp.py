import os, sys, time, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import progan_b200
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
def log(*a):
    print("[r%d %.1f]" % (rank, time.time() % 1000), *a, flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
log("pg ok")
use_graph = "--graph" in sys.argv
K = progan_b200.get_kernels(); K.conv_impl, K.wgrad_tc = "tc", True
torch.manual_seed(0)
G = progan_b200.Generator(128, 128, tanh=False).to(dev); D = progan_b200.Discriminator(128).to(dev)
Gr = progan_b200.Generator(128, 128, tanh=False).to(dev)
for p in list(G.parameters()) + list(D.parameters()) + list(Gr.parameters()):
    dist.broadcast(p.data, 0)
torch.cuda.synchronize(); log("broadcast ok")
tr = progan_b200.Trainer(G, D, Gr, use_graph=use_graph)
res = int(os.environ.get("RES", 32)); step = {8: 1, 16: 2, 32: 3, 64: 4, 128: 5}[res]
g = torch.Generator().manual_seed(1234 + rank)
real = (torch.rand(16, 3, res, res, generator=g) * 2 - 1).to(dev); z = torch.randn(16, 128, generator=g).to(dev); eps = torch.rand(16, 1, 1, 1, generator=g).to(dev)
for i in range(4):
    tr.step(real, z, eps, step, 0.5); torch.cuda.synchronize(); log("step", i, "ok")
t = tr.bD.p.clone(); dist.all_reduce(t); torch.cuda.synchronize()
log("replicas identical:", bool(torch.allclose(t / world, tr.bD.p, rtol=0, atol=1e-6)), tr.read_metrics())
dist.destroy_process_group(); log("done")
