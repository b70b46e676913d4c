"""Generate tests/golden/loss.npz and adam.npz FROM THE REFERENCE'S OWN PYTHON (build container only).

    python tests/golden/make_golden_loss.py

loss.npz   the image term of get_loss (train.py:310,317) evaluated by the reference's own `calc_ssim` / `_ssim` /
           `create_window` / `gaussian` (external.py:71-116) and `l1_loss_v1` (helpers.py:115-116), cut out of their
           source files with `ast` and executed unmodified on the CPU in float32, with torch.autograd gradients
           w.r.t. the rendered image, cam_m and cam_c.
adam.npz   torch.optim.Adam(param_groups, lr=0.0, eps=1e-15) exactly as initialize_optimizer builds it
           (train.py:272-297: one named group per parameter with its own lr), stepped on fixed gradients.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, cut  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    ns = {"torch": torch, "np": np}
    exec("import torch.nn.functional as func\nfrom torch.autograd import Variable\nfrom math import exp\n", ns)
    exec(cut(os.path.join(REF, "external.py"), {"gaussian", "create_window", "calc_ssim", "_ssim"}), ns)
    exec(cut(os.path.join(REF, "helpers.py"), {"l1_loss_v1"}), ns)

    out = {}
    cases = {"a": (40, 37, True), "b": (23, 50, True), "c": (9, 70, False), "d": (64, 64, True)}
    for name, (h, w, affine) in cases.items():
        render = rng.uniform(0, 1, (3, h, w)).astype(np.float32)
        # a target correlated with the render (as in training) plus structure, so SSIM is away from 0
        target = np.clip(render * 0.7 + 0.3 * rng.uniform(0, 1, (3, h, w)), 0, 1).astype(np.float32)
        if name == "d":
            target[:, :16] = render[:, :16]                      # exact ties: sign(0) = 0 in the L1 gradient
        cam_m = rng.normal(0, 0.1, 3).astype(np.float32)
        cam_c = rng.normal(0, 0.05, 3).astype(np.float32)
        r = torch.tensor(render, requires_grad=True)
        t = torch.tensor(target)
        m = torch.tensor(cam_m, requires_grad=True)
        c = torch.tensor(cam_c, requires_grad=True)
        im = torch.exp(m)[:, None, None] * r + c[:, None, None] if affine else r      # train.py:310
        l1 = ns["l1_loss_v1"](im, t)
        ss = ns["calc_ssim"](im, t)
        loss = 0.8 * l1 + 0.2 * (1.0 - ss)                                             # train.py:317
        loss.backward()
        out[name + "_render"], out[name + "_target"] = render, target
        out[name + "_affine"] = np.array(int(affine))
        out[name + "_cam_m"], out[name + "_cam_c"] = cam_m, cam_c
        out[name + "_terms"] = np.array([l1.item(), ss.item(), loss.item()], np.float64)
        out[name + "_d_render"] = r.grad.numpy()
        if affine:
            out[name + "_d_cam_m"], out[name + "_d_cam_c"] = m.grad.numpy(), c.grad.numpy()
    out["window_1d"] = ns["gaussian"](11, 1.5).numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)

    # ---- Adam exactly as initialize_optimizer builds it ----
    lrs = {"means3D": 0.0, "rgb_colors": 0.0025, "unnorm_rotations": 0.001, "log_scales": 0.001, "cam_m": 1e-4}
    shapes = {"means3D": (50, 3), "rgb_colors": (50, 3), "unnorm_rotations": (50, 4), "log_scales": (50, 3), "cam_m": (24, 3)}
    params = {k: torch.nn.Parameter(torch.tensor(rng.normal(0, 1, shapes[k]).astype(np.float32))) for k in lrs}
    param_groups = [{"params": [v], "name": k, "lr": lrs[k]} for k, v in params.items()]
    opt = torch.optim.Adam(param_groups, lr=0.0, eps=1e-15)
    adam = {k + "_init": v.detach().numpy().copy() for k, v in params.items()}
    steps = 7
    for k in lrs:
        adam[k + "_grads"] = (rng.normal(0, 1, (steps,) + shapes[k]) * rng.uniform(1e-6, 1.0, (steps, 1, 1))).astype(np.float32)
    for s in range(steps):
        if s == 4:                                               # update_optimizer (helpers.py:801-804)
            for g in opt.param_groups:
                if g["name"] == "means3D":
                    g["lr"] = 0.000016
        for k, v in params.items():
            v.grad = torch.tensor(adam[k + "_grads"][s])
        opt.step()
        opt.zero_grad(set_to_none=True)
    for k, v in params.items():
        adam[k + "_final"] = v.detach().numpy().copy()
        adam[k + "_exp_avg"] = opt.state[v]["exp_avg"].numpy().copy()
        adam[k + "_exp_avg_sq"] = opt.state[v]["exp_avg_sq"].numpy().copy()
    adam["steps"] = np.array(steps)
    adam["lr_change_step"] = np.array(4)
    np.savez_compressed(os.path.join(HERE, "adam.npz"), **adam)
    print("loss.npz / adam.npz written to", HERE)


if __name__ == "__main__":
    main()
