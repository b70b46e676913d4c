"""Shared checker: an implementation of params2rendervar against the golden vectors the reference's own function
produced (tests/golden/make_golden_activations.py).  Used on the CPU with the reference expression itself (which
validates the checker and the fixture) and on the GPU with the fused kernel (topo4d_b200.activations)."""
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "activations.npz")


def run_and_check(params2rendervar, device):
    g = np.load(G)
    names = ("means3D", "rgb_colors", "unnorm_rotations", "logit_opacities", "log_scales")
    params = {k: torch.tensor(g["in_" + k], device=device, requires_grad=True) for k in names}
    rv = params2rendervar(params)
    assert sorted(rv.keys()) == list(g["keys"])
    assert rv["means3D"] is params["means3D"] and rv["colors_precomp"] is params["rgb_colors"]
    assert rv["means2D"].requires_grad and rv["means2D"].shape == params["means3D"].shape
    assert float(rv["means2D"].detach().abs().max()) == 0.0
    rv["means2D"].retain_grad()                                   # train.py:304 -- must be legal
    loss = sum((rv[k] * torch.tensor(g["w_" + k], device=device)).sum() for k in ("rotations", "opacities", "scales"))
    loss.backward()
    for k in ("rotations", "opacities", "scales"):
        out = rv[k].detach().cpu().numpy()
        assert out.shape == g["out_" + k].shape, k
        np.testing.assert_allclose(out, g["out_" + k], rtol=2e-6, atol=1e-7, err_msg=k)
    for k in ("unnorm_rotations", "logit_opacities", "log_scales"):
        got, ref = params[k].grad.cpu().numpy(), g["grad_" + k]
        assert got.shape == ref.shape, k
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6 * float(np.abs(ref[np.abs(ref) < 1e6]).max()), err_msg=k)
    return rv
