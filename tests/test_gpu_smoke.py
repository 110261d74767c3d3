"""GPU: the driver's smoke() entry point in a FRESH process (default backend flags, nothing set
by other tests), and the module-level drop-in with default settings."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_smoke_entry_point_in_fresh_process():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "--smoke"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "smoke:" in res.stdout


def test_default_modules_train_step_in_fresh_process():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch, progan_b200\n"
        "G = progan_b200.Generator(128, 128, tanh=False).cuda(); D = progan_b200.Discriminator(128).cuda()\n"
        "z = torch.randn(8, 128, device='cuda'); real = torch.rand(8, 3, 32, 32, device='cuda') * 2 - 1\n"
        "fake = G(z, step=3, alpha=0.5)\n"
        "(D(fake.detach(), step=3, alpha=0.5).mean() - D(real, step=3, alpha=0.5).mean()).backward()\n"
        "eps = torch.rand(8, 1, 1, 1, device='cuda')\n"
        "x_hat = (eps * real + (1 - eps) * fake.detach()).requires_grad_(True)\n"
        "g, = torch.autograd.grad(D(x_hat, step=3, alpha=0.5).sum(), x_hat, create_graph=True)\n"
        "gp = 10 * ((g.view(8, -1).norm(2, dim=1) - 1) ** 2).mean(); gp.backward()\n"
        "(-D(fake, step=3, alpha=0.5).mean()).backward()\n"
        "torch.cuda.synchronize()\n"
        "assert all(torch.isfinite(p.grad).all() for p in D.parameters() if p.grad is not None)\n"
        "assert sum(p.grad is not None for p in G.parameters()) > 10\n"
        "print('ok', float(gp))\n" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
