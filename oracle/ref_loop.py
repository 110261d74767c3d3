"""TEST/BENCH INFRASTRUCTURE ONLY — never imported by the product package.

The hot-loop body of the reference's train.py:97-169 (n_critic = 1) driven on the REAL
reference modules (progan_modules.Generator / Discriminator staged under baseline/_ref/ or
imported from /root/reference): same calls in the same order as the script — D.zero_grad,
D(real) with the 0.001 drift term and backward(mone), G(z), D(fake.detach()).mean().backward(one),
x_hat, autograd.grad(create_graph=True), 10 * mean((||g|| - 1)^2).backward(), d_optimizer.step(),
G.zero_grad / D.zero_grad, -D(fake).mean().backward(), g_optimizer.step(), accumulate().
Data loading, tqdm, sample grids and checkpoints of the script are left out (the
script itself no longer runs on current torchvision)."""
import os
import sys

import torch


def import_reference():
    """The reference's progan_modules: staged copy first (GPU box), live checkout second."""
    from oracle import stage_reference
    d = stage_reference.staged_dir() or ("/root/reference" if os.path.isdir("/root/reference") else None)
    if d is None:
        return None, None
    if d not in sys.path:
        sys.path.insert(0, d)
    import progan_modules as R
    return R, d


def accumulate(model1, model2, decay=0.999):
    """train.py:17-22."""
    par1 = dict(model1.named_parameters())
    par2 = dict(model2.named_parameters())
    for k in par1.keys():
        par1[k].data.mul_(decay).add_(par2[k].data, alpha=1 - decay)


def iteration(G, D, Grun, g_opt, d_opt, real, z, eps, step, alpha):
    """One iteration of train.py:97-169; returns (disc_loss, grad_penalty, gen_loss) tensors."""
    one = torch.tensor(1, dtype=torch.float, device=real.device)
    mone = one * -1
    D.zero_grad()                                                          # :98
    real_predict = D(real, step=step, alpha=alpha)                         # :126
    real_predict = real_predict.mean() - 0.001 * (real_predict ** 2).mean()   # :128-129
    real_predict.backward(mone)                                            # :130
    fake_image = G(z, step=step, alpha=alpha)                              # :135
    fake_predict = D(fake_image.detach(), step=step, alpha=alpha)          # :136
    fake_predict = fake_predict.mean()
    fake_predict.backward(one)                                             # :139
    x_hat = eps * real.data + (1 - eps) * fake_image.detach().data         # :143
    x_hat.requires_grad = True
    hat_predict = D(x_hat, step=step, alpha=alpha)
    grad_x_hat = torch.autograd.grad(outputs=hat_predict.sum(), inputs=x_hat, create_graph=True)[0]
    grad_penalty = ((grad_x_hat.view(grad_x_hat.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()
    grad_penalty = 10 * grad_penalty
    grad_penalty.backward()                                                # :151
    d_opt.step()                                                           # :155
    G.zero_grad()
    D.zero_grad()
    predict = D(fake_image, step=step, alpha=alpha)                        # :162
    loss = -predict.mean()
    loss.backward()                                                        # :167
    g_opt.step()
    accumulate(Grun, G)                                                    # :169
    return (real_predict - fake_predict).detach(), grad_penalty.detach(), loss.detach()


def build(R, channel=128, zdim=128, device="cpu"):
    """Models and optimisers exactly as train.py:243-259 builds them."""
    from torch import optim
    G = R.Generator(in_channel=channel, input_code_dim=zdim, pixel_norm=True, tanh=False).to(device)
    D = R.Discriminator(feat_dim=channel).to(device)
    Grun = R.Generator(in_channel=channel, input_code_dim=zdim, pixel_norm=True, tanh=False).to(device)
    Grun.train(False)
    g_opt = optim.Adam(G.parameters(), lr=0.001, betas=(0.0, 0.99))
    d_opt = optim.Adam(D.parameters(), lr=0.001, betas=(0.0, 0.99))
    accumulate(Grun, G, 0)
    return G, D, Grun, g_opt, d_opt
