"""Generate the golden fixtures from the REAL reference (run in the build container only:
needs /root/reference).  Usage:  python tests/golden/make_golden.py

For every case of common.CASES it loads seeded weights into the reference's own
Generator/Discriminator (progan_modules.py), executes the hot-loop body of
train.py:122-169 verbatim in spirit (same calls, same order, torch.optim.Adam with
betas (0.0, 0.99)), and stores outputs + compact gradient summaries in <case>.pt.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402

REF = "/root/reference"


def run_case(name):
    sys.path.insert(0, REF)
    import progan_modules as R
    from torch import optim
    from torch.autograd import grad

    inp = common.make_inputs(name)
    step, alpha = inp["step"], inp["alpha"]
    G, D = common.build(R, name, inp)
    Grun, _ = common.build(R, name, inp)
    lab = inp["label"]
    la = (lab,) if lab is not None else ()          # conditional models take the label second
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in inp["G"].items()}
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in inp["D"].items()}
    G.load_state_dict(inp["G"]); D.load_state_dict(inp["D"]); Grun.load_state_dict(inp["G"])
    g_opt = optim.Adam(G.parameters(), lr=0.001, betas=(0.0, 0.99))
    d_opt = optim.Adam(D.parameters(), lr=0.001, betas=(0.0, 0.99))
    real, z, eps = inp["real"], inp["z"], inp["eps"]
    one = torch.tensor(1, dtype=torch.float)
    mone = one * -1
    out = {}

    # --- train.py:98, 122-155
    D.zero_grad()
    b_size = real.size(0)
    real_predict_raw = D(real, *la, step=step, alpha=alpha)
    real_predict = real_predict_raw.mean() - 0.001 * (real_predict_raw ** 2).mean()
    real_predict.backward(mone)
    fake_image = G(z, *la, step=step, alpha=alpha)
    fake_predict = D(fake_image.detach(), *la, step=step, alpha=alpha)
    fake_predict = fake_predict.mean()
    fake_predict.backward(one)
    x_hat = eps * real.data + (1 - eps) * fake_image.detach().data
    x_hat.requires_grad = True
    hat_predict = D(x_hat, *la, step=step, alpha=alpha)
    grad_x_hat = grad(outputs=hat_predict.sum(), inputs=x_hat, create_graph=True)[0]
    grad_penalty = ((grad_x_hat.view(grad_x_hat.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()
    grad_penalty = 10 * grad_penalty
    grad_penalty.backward()
    out["real_predict"] = real_predict_raw.detach().clone()
    out["fake"] = fake_image.detach().clone()
    out["hat_predict"] = hat_predict.detach().clone()
    out["grad_x_hat"] = grad_x_hat.detach().clone()
    out["grad_penalty"] = grad_penalty.detach().clone()
    out["disc_loss"] = (real_predict - fake_predict).detach().clone()
    out["d_grads"] = common.summarize_dict({k: p.grad for k, p in D.named_parameters()
                                            if p.grad is not None})
    out["d_grad_none"] = sorted(k for k, p in D.named_parameters() if p.grad is None)
    d_opt.step()
    out["d_params_after"] = common.summarize_dict(dict(D.named_parameters()))
    # --- train.py:158-169
    G.zero_grad(); D.zero_grad()
    predict = D(fake_image, *la, step=step, alpha=alpha)
    loss = -predict.mean()
    loss.backward()
    out["gen_loss"] = loss.detach().clone()
    out["g_grads"] = common.summarize_dict({k: p.grad for k, p in G.named_parameters()
                                            if p.grad is not None})
    out["g_grad_none"] = sorted(k for k, p in G.named_parameters() if p.grad is None)
    g_opt.step()
    out["g_params_after"] = common.summarize_dict(dict(G.named_parameters()))
    # accumulate(g_running, generator)  train.py:17-22 (decay 0.999)
    par1, par2 = dict(Grun.named_parameters()), dict(G.named_parameters())
    for k in par1:
        par1[k].data.mul_(0.999).add_(par2[k].data, alpha=1 - 0.999)
    out["g_running_after"] = common.summarize_dict(par1)
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name in common.ALL_CASES:
        if only and name not in only and common.family(name) not in only:
            continue
        res = run_case(name)
        path = os.path.join(common.HERE, name + ".pt")
        torch.save(res, path)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB",
              "gp=%.6f" % float(res["grad_penalty"]))
