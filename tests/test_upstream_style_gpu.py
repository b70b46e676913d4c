"""The upstream-style CUDA arm (tools/upstream_style, bench.py --impl upstream_style) is only a benchmark denominator, but a
denominator that computes something else would be worthless: its images must agree with the oracle's and its gradients with the
oracle-verified product path on the same inputs."""
import numpy as np
import pytest
import torch

from tests import parity
from topo4d_b200 import engine, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("deg", [None, 2])
def test_upstream_style_arm_matches_oracle_and_product(deg):
    from tools import upstream_style as US
    dev = torch.device("cuda:0")
    scene = synth.random_scene(5000, seed=3, sh_degree=deg)
    cam = synth.make_camera(synth.look_at((1.0, 0.5, -3.5)), 256, 208, 256.0, 256.0)
    H, W = 208, 256
    bg = (0.2, 0.4, 0.6)
    t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
    cam_t = torch.tensor(engine.pack_cameras_numpy([cam], bg), device=dev)
    color, radii, depth, alpha, view = US.forward(t, cam_t, H, W, deg or 0, bg)
    ref = parity.oracle_forward(scene, [cam], H, W, deg or 0, bg)[0]
    assert US.num_rendered() == ref["state"].num_rendered
    err = np.abs(color.cpu().numpy() - ref["color"]).max(0)
    # expf vs the oracle's exp: a handful of pixels may flip a discrete alpha / T threshold
    assert (err > parity.ABS_TOL).sum() <= 5, int((err > parity.ABS_TOL).sum())
    assert np.abs(depth.cpu().numpy() - ref["depth"]).max() < 1e-2 and np.abs(alpha.cpu().numpy() - ref["alpha"]).max() < 5e-3
    # gradients vs the product path (itself checked against the oracle)
    gen = torch.Generator(device=dev).manual_seed(5)
    gC, gD, gA = (torch.randn(s, device=dev, generator=gen) for s in ((3, H, W), (1, H, W), (1, H, W)))
    M = 0 if deg is None else (deg + 1) ** 2
    _, n = engine.flat_layout(5000, M, deg is not None, False)
    seg = US.backward(view, gC, gD, gA, torch.empty(n, device=dev), bg)
    ours = engine.forward(t["means3D"], t["opacities"], cam_t, H, W, shs=t.get("shs"), colors_precomp=t.get("colors_precomp"),
                          scales=t["scales"], rotations=t["rotations"], sh_degree=deg or 0)
    gb = engine.backward(ours[4], gC[None], gD[None], gA[None])
    for k, v in seg.items():
        b = getattr(gb, k)
        rel = parity.grad_rel_err(v.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64).reshape(v.shape))
        assert rel < 5e-3, (k, rel)
