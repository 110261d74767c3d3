"""Shared test helpers: build the product modules from a state dict, run the train-step
sequence of train.py:122-167 through the product's public API, relative errors."""
import torch

import progan_b200
from progan_b200 import functions as F_


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_models(inp, precision, device="cpu", dtype=torch.float32, name=None):
    import common
    if name:
        G, D = common.build(progan_b200, name, inp, precision=precision)
    else:
        G = progan_b200.Generator(input_code_dim=inp["z_dim"], in_channel=inp["channel"],
                                  pixel_norm=inp["pixel_norm"], tanh=inp["tanh"], precision=precision)
        D = progan_b200.Discriminator(feat_dim=inp["channel"], precision=precision)
    G.load_state_dict(inp["G"])
    D.load_state_dict(inp["D"])
    return G.to(device=device, dtype=dtype), D.to(device=device, dtype=dtype)


def product_train_step(G, D, real, z, eps, step, alpha, fused_gp=True, label=None):
    """The loop body of train.py:122-151 + 158-167 written against the product API (the same
    calls the reference script makes; fused_gp swaps the torch norm chain for the GP kernel)."""
    D.zero_grad(set_to_none=True)
    G.zero_grad(set_to_none=True)
    b = real.size(0)
    la = (label,) if label is not None else ()
    real_raw = D(real, *la, step=step, alpha=alpha)
    real_predict = real_raw.mean() - 0.001 * (real_raw ** 2).mean()
    (-real_predict).backward()
    fake = G(z, *la, step=step, alpha=alpha)
    fake_predict = D(fake.detach(), *la, step=step, alpha=alpha).mean()
    fake_predict.backward()
    x_hat = (eps * real.data + (1 - eps) * fake.detach().data).requires_grad_(True)
    hat = D(x_hat, *la, step=step, alpha=alpha)
    (g,) = torch.autograd.grad(outputs=hat.sum(), inputs=x_hat, create_graph=True)
    if fused_gp:
        gp = F_.gradient_penalty(g, 10.0)
    else:
        gp = 10 * ((g.view(b, -1).norm(2, dim=1) - 1) ** 2).mean()
    gp.backward()
    res = dict(real_predict=real_raw.detach(), fake=fake.detach(), hat_predict=hat.detach(),
               grad_x_hat=g.detach(), grad_penalty=gp.detach(),
               disc_loss=(real_predict - fake_predict).detach(),
               d_grads={k: p.grad.clone() for k, p in D.named_parameters() if p.grad is not None})
    # G phase (same D weights here; tests that include the optimiser do the update in between)
    return res, fake


def product_g_phase(G, D, fake, step, alpha, label=None):
    G.zero_grad(set_to_none=True)
    D.zero_grad(set_to_none=True)
    la = (label,) if label is not None else ()
    loss = -D(fake, *la, step=step, alpha=alpha).mean()
    loss.backward(inputs=list(G.parameters()))
    return loss.detach(), {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
