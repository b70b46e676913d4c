"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle, same seeded inputs."""
import numpy as np
import pytest
import torch

from tests import parity
from topo4d_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _scaled(scene, k):
    scene = dict(scene)
    scene["scales"] = scene["scales"] * k
    return scene


@pytest.mark.parametrize("bg", [(0.0, 0.0, 0.0), (0.2, 0.5, 0.8)])
def test_config1_colors_precomp(bg):
    """BASELINE config 1: 1 view 256x256, 5k random Gaussians, colors_precomp, fwd+bwd."""
    scene = synth.random_scene(5000, seed=0)
    m = parity.compare(scene, [synth.front_camera(256, 256)], 256, 256, 0, bg)
    parity.assert_parity(m, allow_flips=2)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_config1_sh(deg):
    scene = synth.random_scene(5000, seed=deg + 10, sh_degree=deg)
    cam = synth.make_camera(synth.look_at((1.5, 0.8, -3.5)), 256, 256, 256.0, 256.0)
    m = parity.compare(scene, [cam], 256, 256, deg, (0.1, 0.1, 0.1))
    parity.assert_parity(m, allow_flips=2)


def test_sh_more_coeffs_than_degree():
    """shs [N,16,3] rendered at sh_degree 1: unused bands get zero gradient."""
    scene = synth.random_scene(2000, seed=3, sh_degree=3)
    m = parity.compare(scene, [synth.front_camera(128, 96)], 96, 128, 1)
    parity.assert_parity(m, allow_flips=2)


def test_ragged_image_and_big_splats():
    """H, W not multiples of 16; splats spanning many tiles; long per-tile lists (> 1 chunk)."""
    scene = _scaled(synth.random_scene(3000, seed=4), 4.0)
    cam = synth.make_camera(synth.look_at((0.0, 0.0, -3.0)), 200, 137, 180.0, 170.0, 90.0, 70.0)
    m = parity.compare(scene, [cam], 137, 200, 0, (0.3, 0.2, 0.1))
    assert m["num_rendered"] > 50000
    parity.assert_parity(m, allow_flips=2)


def test_multi_view_batch_sums_gradients():
    """V = 3 views in one call == sum over three single-view oracle passes."""
    scene = synth.random_scene(3000, seed=5, sh_degree=2)
    cams = [synth.make_camera(synth.look_at(e), 160, 128, 200.0, 200.0) for e in ((0, 0, -4.0), (3.0, 1.0, -2.5), (-2.0, -1.5, 3.0))]
    m = parity.compare(scene, cams, 128, 160, 2, (0.0, 0.0, 0.0))
    parity.assert_parity(m, allow_flips=3)


def test_cov3d_precomp_and_scale_modifier():
    scene = synth.random_scene(1500, seed=6)
    # build cov3D from scale/rot on the host (R S S^T R^T, external.py:26-43 convention)
    q, s = scene["rotations"].astype(np.float64), scene["scales"].astype(np.float64) * 0.7
    r, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    Sig = np.einsum("nij,nj,nkj->nik", R, s * s, R)
    cov = np.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1).astype(np.float32)
    sc2 = {k: v for k, v in scene.items() if k not in ("scales", "rotations")}
    sc2["cov3D_precomp"] = cov
    m = parity.compare(sc2, [synth.front_camera(128, 128)], 128, 128, 0)
    parity.assert_parity(m, allow_flips=2)
    m = parity.compare(scene, [synth.front_camera(128, 128)], 128, 128, 0, scale_modifier=0.7)
    parity.assert_parity(m, allow_flips=2)


def test_topo4d_regime_opacity_one_mesh_bound():
    """Reference regime: mesh-bound isotropic splats, opacity = sigmoid(1000) = 1 (train.py:142), so the
    0.99 cap is active at every centre and the backward must be straight-through."""
    scene = synth.head_scene(8280, seed=0, sh_degree=None, opacity="topo4d")
    cams = synth.ring_cameras(2, w=512, h=375, radius=0.6, focal_over_h=1.6)
    m = parity.compare(scene, cams, 375, 512, 0, noise_floor=True)
    parity.assert_parity(m, allow_flips=3)


def test_empty_and_culled_scenes():
    dev = torch.device("cuda:0")
    cam = torch.tensor(engine.pack_cameras_numpy([synth.front_camera(64, 48)], (0.5, 0.5, 0.5)), device=dev)
    # N = 0 -> zero image, not bg
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, depth, alpha, st = engine.forward(z(0, 3), z(0, 1), cam, 48, 64, colors_precomp=z(0, 3), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (1, 3, 48, 64) and float(color.abs().max()) == 0.0 and radii.shape == (1, 0)
    # everything behind the camera -> bg everywhere, radii 0, zero grads
    sc = synth.random_scene(100, seed=1)
    sc["means3D"][:, 2] -= 50.0
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    color, radii, depth, alpha, st = engine.forward(t["means3D"], t["opacities"], cam, 48, 64, colors_precomp=t["colors_precomp"],
                                                    scales=t["scales"], rotations=t["rotations"])
    assert int(radii.abs().sum()) == 0 and st.status().num_instances == 0
    assert torch.allclose(color, torch.full_like(color, 0.5)) and float(alpha.max()) == 0.0
    g = engine.backward(st, torch.ones_like(color), torch.ones_like(depth), torch.ones_like(alpha))
    assert float(g.flat.abs().max()) == 0.0


def test_capacity_overflow_is_detected_and_recovered():
    dev = torch.device("cuda:0")
    sc = synth.random_scene(2000, seed=2)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    cam = torch.tensor(engine.pack_cameras_numpy([synth.front_camera(128, 128)]), device=dev)
    kw = dict(colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"])
    ref = engine.forward(t["means3D"], t["opacities"], cam, 128, 128, **kw)
    need = ref[4].status().num_instances
    # explicit too-small capacity, no sync: status must flag overflow, nothing may be written out of bounds
    bad = engine.forward(t["means3D"], t["opacities"], cam, 128, 128, check="none", cap_instances=need // 3, **kw)
    st = bad[4].status()
    assert st.overflow == 1 and st.num_instances == need
    # sync mode grows and re-runs transparently
    good = engine.forward(t["means3D"], t["opacities"], cam, 128, 128, check="sync", cap_instances=need // 3, **kw)
    assert good[4].status().overflow == 0
    assert torch.equal(good[0], ref[0])


def test_forward_is_deterministic_and_backward_close():
    dev = torch.device("cuda:0")
    sc = synth.random_scene(4000, seed=9)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    cam = torch.tensor(engine.pack_cameras_numpy([synth.front_camera(192, 160)]), device=dev)
    kw = dict(colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"])
    a = engine.forward(t["means3D"], t["opacities"], cam, 160, 192, **kw)
    b = engine.forward(t["means3D"], t["opacities"], cam, 160, 192, **kw)
    for k in range(4):
        assert torch.equal(a[k], b[k])                       # bit-identical images / radii
    assert torch.equal(a[4].view()["sorted_ids"], b[4].view()["sorted_ids"])
    gc = torch.randn_like(a[0])
    ga = engine.backward(a[4], gc).flat
    gb = engine.backward(b[4], gc).flat
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-5 * float(ga.abs().max()))   # atomics reorder fp32 sums


def test_blend_px_variants_agree():
    """The pixels-per-thread tuning hint (4 / 2 / 1; 8 is accepted and means 4) never changes results: forward bit-identical, backward equal
    up to the order of fp32 partial sums."""
    dev = torch.device("cuda:0")
    sc = synth.random_scene(4000, seed=21, sh_degree=1)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    cams = [synth.front_camera(200, 137), synth.make_camera(synth.look_at((2.0, 1.0, -3.0)), 200, 137, 150.0, 150.0)]
    cam = torch.tensor(engine.pack_cameras_numpy(cams, (0.2, 0.1, 0.3)), device=dev)
    kw = dict(shs=t["shs"], scales=t["scales"] * 2.0, rotations=t["rotations"], sh_degree=1)
    ref = engine.forward(t["means3D"], t["opacities"], cam, 137, 200, blend_px=4, **kw)
    gc, gd, ga = torch.randn_like(ref[0]), torch.randn_like(ref[2]), torch.randn_like(ref[3])
    gref = engine.backward(ref[4], gc, gd, ga).flat
    for px in (8, 2, 1):
        out = engine.forward(t["means3D"], t["opacities"], cam, 137, 200, blend_px=px, **kw)
        for k in range(4):
            assert torch.equal(out[k], ref[k]), (px, k)
        v0, v1 = ref[4].view(), out[4].view()
        assert torch.equal(v0["n_contrib"], v1["n_contrib"]) and torch.equal(v0["final_T"], v1["final_T"])
        g = engine.backward(out[4], gc, gd, ga).flat
        assert torch.allclose(g, gref, rtol=2e-4, atol=2e-5 * float(gref.abs().max())), px
    assert engine.pick_blend_px(None) == 4 and engine.pick_blend_px(40000) == 4 and engine.pick_blend_px(5000) == 2 and engine.pick_blend_px(300) == 1


@pytest.mark.parametrize("n", [17000, 5200, 1500, 700, 300])
def test_very_long_tile_lists_use_the_global_sort_path(n):
    """Every tile list has exactly n entries: n = 17000 is longer than the 16384-key shared-memory sort of the long-list kernel
    (in-place global bitonic network), 5200 takes that kernel (dozens of record chunks per tile, early termination deep
    inside the list); 1500 / 700 / 300 take the 16 / 8 / 4 keys-per-thread register networks with +inf padding."""
    rng = np.random.default_rng(31)
    scene = dict(means3D=(rng.normal(0, 0.05, (n, 3))).astype(np.float32),
                 scales=np.full((n, 3), 0.6, np.float32) * rng.uniform(0.8, 1.2, (n, 3)).astype(np.float32),
                 rotations=np.tile(np.array([[1, 0, 0, 0]], np.float32), (n, 1)),
                 opacities=rng.uniform(0.004, 0.05, (n, 1)).astype(np.float32),
                 colors_precomp=rng.uniform(0, 1, (n, 3)).astype(np.float32))
    cam = synth.front_camera(48, 40, dist=4.0, fx=48.0)
    # 17000-deep lists: forward and bit-exact sorted lists only (replaying 17000 layers back to front in fp32 is beyond the
    # 1e-3 gradient bound whatever the sort did; the backward at depth is covered by n = 5200)
    m = parity.compare(scene, [cam], 40, 48, 0, (0.1, 0.2, 0.3), with_backward=n <= 6000)
    assert m["num_rendered"] == n * 9          # 3 x 3 tiles, every splat covers them all
    parity.assert_parity(m, allow_flips=2)
    if n == 5200:
        # the list-length hint never changes results: the second call of the shape knows the lists are long (long-list kernel),
        # and a caller that wrongly promises short lists still gets the in-place global network
        m = parity.compare(scene, [cam], 40, 48, 0, (0.1, 0.2, 0.3))
        parity.assert_parity(m, allow_flips=2)
        for k in list(engine._MAXTILE_MEMO):
            engine._MAXTILE_MEMO[k] = 100
        m = parity.compare(scene, [cam], 40, 48, 0, (0.1, 0.2, 0.3))
        parity.assert_parity(m, allow_flips=2)
        engine._MAXTILE_MEMO.clear()


def test_odd_width_and_degenerate_inputs():
    """W % 4 != 0 (scalar store path), zero / negative opacity, a splat with a NaN mean (culled), huge splat."""
    scene = synth.random_scene(1500, seed=33)
    scene["opacities"][:50] = 0.0
    scene["opacities"][50:60] = -0.5
    scene["means3D"][60] = np.nan
    scene["scales"][61] = 50.0
    cam = synth.make_camera(synth.look_at((0.5, 0.2, -3.0)), 131, 77, 120.0, 110.0)
    m = parity.compare(scene, [cam], 77, 131, 0, (0.3, 0.3, 0.3))
    parity.assert_parity(m, allow_flips=2)


def test_repeated_backward_and_staged_forward_reuse_the_work_queues():
    """The device-side work queues rewind themselves (no memset in the stream): a second backward on the same state,
    and a forward issued one stage per call, give the results of the plain path."""
    dev = torch.device("cuda:0")
    sc = synth.random_scene(3000, seed=41)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    cam = torch.tensor(engine.pack_cameras_numpy([synth.front_camera(160, 128)]), device=dev)
    kw = dict(colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"])
    a = engine.forward(t["means3D"], t["opacities"], cam, 128, 160, **kw)
    b = engine.forward(t["means3D"], t["opacities"], cam, 128, 160, stage_events={}, **kw)
    for k in range(4):
        assert torch.equal(a[k], b[k])
    gc = torch.randn_like(a[0])
    g1 = engine.backward(a[4], gc).flat.clone()
    g2 = engine.backward(a[4], gc).flat.clone()
    g3 = engine.backward(b[4], gc, stage_events={}).flat.clone()
    tol = 1e-5 * float(g1.abs().max())
    assert float(g1.abs().max()) > 0
    assert torch.allclose(g1, g2, rtol=1e-4, atol=tol) and torch.allclose(g1, g3, rtol=1e-4, atol=tol)


# ---- BASELINE config 2 at FULL size: the workload bench.py times (60 k mesh-bound Gaussians, SH degree 3, 1920x1080) ----
_CFG2_VIEWS = (0, 7, 13, 20)          # both camera rings, four azimuths
# Flip budget at this size, stated: at most 1 pixel per 100 000 may differ in a discrete threshold decision
# (alpha < 1/255, T(1-alpha) < 1e-4 -- a 1-ulp exp() event); flipped pixels are counted, excluded from the value
# comparison and their loss gradients zeroed on both sides (tests/parity.py).  4 views x 2 073 600 px -> <= 82.
_CFG2_FLIPS = len(_CFG2_VIEWS) * 1920 * 1080 // 100000


def _cfg2(opacity):
    scene = synth.head_scene(60000, seed=0, sh_degree=3, opacity=opacity)
    cams = [synth.ring_cameras(24)[i] for i in _CFG2_VIEWS]
    return scene, cams


@pytest.mark.parametrize("opacity", ["topo4d", "generic"])
@pytest.mark.parametrize("px", [2, 4, 1])
def test_config2_full_size_parity(opacity, px):
    """SURVEY 8(d) configs 2-3, both opacity regimes ("topo4d" = 1.0, "generic" = U(0.3, 1)), with the blend kernels'
    pixels-per-thread variant forced to each of 1 / 2 / 4 (bench.py's 24-view launch runs PX = 2).  Tolerances: RGB /
    depth / alpha <= 1e-4 abs, gradients <= 1e-3 rel (or the all-fp32 algorithm's own noise floor where the 0.99 cap makes
    that larger), radii / tile ranges / sorted lists bit-exact, flips <= _CFG2_FLIPS."""
    scene, cams = _cfg2(opacity)
    m = parity.compare(scene, cams, 1080, 1920, 3, (0.0, 0.0, 0.0), noise_floor=True, blend_px=px, cache_key="cfg2-" + opacity)
    assert m["num_rendered"] > 400000
    parity.assert_parity(m, allow_flips=_CFG2_FLIPS)


def test_config2_full_size_background_and_auto_px():
    """Same scene, non-zero background, blend_px chosen by the engine's heuristic (second call of the shape)."""
    scene, cams = _cfg2("generic")
    for _ in range(2):
        m = parity.compare(scene, cams[:2], 1080, 1920, 3, (0.2, 0.5, 0.8), cache_key="cfg2-generic-bg")
        parity.assert_parity(m, allow_flips=_CFG2_FLIPS)
