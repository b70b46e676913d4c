"""GPU: the reference-facing Python surface (module name, kwargs, autograd contract)."""
import numpy as np
import pytest
import torch

from tests import parity
from topo4d_b200 import synth

pytestmark = pytest.mark.gpu


def _settings(cam, bg, deg=0):
    from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera
    dev = "cuda"
    w2c = torch.tensor(cam.w2c, dtype=torch.float32, device=dev)
    # exactly how the reference hands matrices over: non-contiguous transposed views (helpers.py:67-72)
    view = w2c.unsqueeze(0).transpose(1, 2)
    proj = torch.tensor(cam.projmatrix, device=dev).unsqueeze(0)
    return Camera(image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                  bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=1.0, viewmatrix=view,
                  projmatrix=proj, sh_degree=deg, campos=torch.tensor(cam.campos, device=dev), prefiltered=False, debug=False)


def test_reference_call_pattern_train_py_307():
    """im, radius, _, _ = Renderer(raster_settings=cam)(**rendervar); loss.backward(); means2D.grad retained."""
    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    sc = synth.random_scene(3000, seed=0)
    cam = synth.front_camera(160, 120)
    bg = (0.0, 0.0, 0.0)
    params = {k: torch.tensor(v, device="cuda", requires_grad=True) for k, v in sc.items()}
    rendervar = {
        "means3D": params["means3D"], "colors_precomp": params["colors_precomp"],
        "rotations": torch.nn.functional.normalize(params["rotations"]), "opacities": params["opacities"],
        "scales": params["scales"],
        "means2D": torch.zeros_like(params["means3D"], requires_grad=True, device="cuda") + 0,
    }
    rendervar["means2D"].retain_grad()
    im, radius, depth, alpha = Renderer(raster_settings=_settings(cam, bg))(**rendervar)
    assert im.shape == (3, 120, 160) and radius.shape == (3000,) and radius.dtype == torch.int32
    assert depth.shape == (1, 120, 160) and alpha.shape == (1, 120, 160)
    rng = np.random.default_rng(1)
    gC = rng.normal(size=(1, 3, 120, 160)).astype(np.float32)
    (im * torch.tensor(gC[0], device="cuda")).sum().backward()          # depth/alpha unused -> zero grads, like train.py
    sc_o = dict(sc, rotations=rendervar["rotations"].detach().cpu().numpy())   # what the op actually received
    ref, ref_g = parity.run_oracle(sc_o, [cam], 120, 160, 0, bg, gC, None, None)
    assert np.abs(im.detach().cpu().numpy() - ref[0]["color"]).max() <= parity.ABS_TOL
    assert (radius.cpu().numpy() == ref[0]["radii"]).all()
    assert parity.grad_rel_err(rendervar["means2D"].grad.cpu().numpy(), ref_g["means2D"]) <= parity.REL_TOL
    assert float(rendervar["means2D"].grad[:, 2].abs().max()) == 0.0
    for k in ("means3D", "colors_precomp", "opacities", "scales"):
        assert parity.grad_rel_err(params[k].grad.cpu().numpy().reshape(ref_g[k].shape), ref_g[k]) <= parity.REL_TOL, k
    seen = radius > 0
    assert bool(seen.any())


def test_sh_path_markvisible_and_errors():
    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    sc = synth.random_scene(1000, seed=2, sh_degree=3)
    cam = synth.make_camera(synth.look_at((2.0, 0.5, -3.0)), 96, 96, 100.0, 100.0)
    t = {k: torch.tensor(v, device="cuda") for k, v in sc.items()}
    r = Renderer(raster_settings=_settings(cam, (0.1, 0.2, 0.3), deg=3))
    im, radius, _, _ = r(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]), opacities=t["opacities"],
                         shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    ref, _ = parity.run_oracle(sc, [cam], 96, 96, 3, (0.1, 0.2, 0.3))
    assert np.abs(im.cpu().numpy() - ref[0]["color"]).max() <= parity.ABS_TOL
    from oracle import gs_oracle
    vis = r.markVisible(t["means3D"])
    assert vis.dtype == torch.bool and (vis.cpu().numpy() == gs_oracle.mark_visible(sc["means3D"], cam.viewmatrix)).all()
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=t["means3D"], means2D=None, opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=t["means3D"], means2D=None, opacities=t["opacities"], shs=t["shs"], scales=t["scales"])


def test_cpu_tensors_fail_loudly():
    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    sc = synth.random_scene(10, seed=0)
    t = {k: torch.tensor(v) for k, v in sc.items()}
    cam = synth.front_camera(32, 32)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        Renderer(raster_settings=_settings(cam, (0, 0, 0)))(means3D=t["means3D"], means2D=None, opacities=t["opacities"],
                                                            colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"])


def test_async_mode_defers_the_overflow_check(monkeypatch):
    """TOPO4D_B200_SYNC=0: no host sync per forward; an overflow of call k surfaces at call k+1 and the retry works."""
    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    from topo4d_b200 import engine, rasterizer
    monkeypatch.setenv("TOPO4D_B200_SYNC", "0")
    rasterizer._PENDING.clear()
    sc = synth.random_scene(1777, seed=5)
    cam = synth.front_camera(112, 80)
    t = {k: torch.tensor(v, device="cuda") for k, v in sc.items()}
    r = Renderer(raster_settings=_settings(cam, (0, 0, 0)))

    def call(scale):
        return r(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]), opacities=t["opacities"],
                 colors_precomp=t["colors_precomp"], scales=t["scales"] * scale, rotations=t["rotations"])
    small = call(0.2)[0].clone()                       # sizes the workspace for few instances (first call counts exactly)
    call(6.0)                                          # needs far more instances than the remembered capacity: overflows silently
    with pytest.raises(RuntimeError, match="capacity has been raised"):
        call(0.2)
    big = call(6.0)[0]                                 # retry with the raised capacity
    call(0.2)                                          # would raise if `big` had overflowed again
    ref, _ = parity.run_oracle(dict(sc, scales=sc["scales"] * 6.0), [cam], 80, 112, 0, (0, 0, 0))
    assert np.abs(big.cpu().numpy() - ref[0]["color"]).max() <= parity.ABS_TOL
    ref_s, _ = parity.run_oracle(dict(sc, scales=sc["scales"] * 0.2), [cam], 80, 112, 0, (0, 0, 0))
    assert np.abs(small.cpu().numpy() - ref_s[0]["color"]).max() <= parity.ABS_TOL
    rasterizer._PENDING.clear()


def test_inplace_edit_between_forward_and_backward_raises():
    """The kernels re-read the inputs in backward (nothing but pointers is stored), so an in-place edit of an input
    between forward and backward must raise autograd's version-counter error -- as it does with upstream's wrapper,
    which saves its inputs with save_for_backward -- instead of silently producing gradients of the modified values."""
    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    sc = synth.random_scene(500, seed=4)
    cam = synth.front_camera(64, 48)
    p = {k: torch.tensor(v, device="cuda", requires_grad=True) for k, v in sc.items()}
    scales = p["scales"] * 1.0                       # non-leaf, so an in-place edit is legal for autograd itself
    im, *_ = Renderer(raster_settings=_settings(cam, (0.0, 0.0, 0.0)))(
        means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]), colors_precomp=p["colors_precomp"],
        rotations=p["rotations"], opacities=p["opacities"], scales=scales)
    scales.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        im.sum().backward()
