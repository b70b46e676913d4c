"""GPU: face3d render_colors CUDA kernel vs the reference's own C++ (oracle/_ref, built from the
reference sources) or, where that build is absent, our port of it -- bit-exact coverage and colours."""
import os

import numpy as np
import pytest
import torch

from oracle import f3d_oracle
from topo4d_b200 import synth
from topo4d_b200.face3d_compat import mesh_core_cython as mcc
from topo4d_b200.face3d_compat import render as f3d_render

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
_cpu = f3d_oracle.render_colors_ref if f3d_oracle.have_ref() else f3d_oracle.render_colors_port


def test_golden_fixtures_bit_exact():
    g = np.load(os.path.join(G, "face3d_small.npz"))
    for n in sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_image")}):
        h, w = (int(x) for x in g[n + "_hw"])
        img = f3d_render.render_colors(g[n + "_vertices"], g[n + "_triangles"], g[n + "_colors"], h, w, 3)
        np.testing.assert_array_equal(img, g[n + "_image"], err_msg=n)


# the last six cases walk the stage-1 warp split (16 / 4 / 2 warps per group of 32 triangles, chosen from the mean box size), groups that
# are not full, the flat fast path next to the general one, and widths that are not a multiple of four (scalar shade pass)
@pytest.mark.parametrize("grid,res,zmode", [(31, 256, "flat"), (64, 512, "random"), (300, 1024, "flat"), (7, 333, "random"),
                                            (9, 256, "flat"), (12, 256, "flat"), (20, 256, "random"), (20, 258, "flat"),
                                            (1, 100, "flat"), (5, 1001, "flat")])
def test_matches_reference_cpu(grid, res, zmode):
    v, t, c = synth.uv_grid_mesh(grid=grid, res=res, seed=grid)
    if zmode == "random":
        v[:, 2] = np.random.default_rng(grid).normal(size=v.shape[0])
    ref_img, ref_dep = _cpu(v, t, c, res, res, 3)
    image = np.zeros((res, res, 3), np.float32)
    depth = np.zeros((res, res), np.float32) - 999999.0
    ret = mcc.render_colors_core(image, v.astype(np.float32), t.astype(np.int32), c.astype(np.float32), depth,
                                 v.shape[0], t.shape[0], res, res, 3)
    assert ret is None
    np.testing.assert_array_equal(image != 0, ref_img != 0)          # coverage bit-exact
    np.testing.assert_array_equal(image, ref_img)                    # colours bit-exact (same unfused fp32 ops)
    np.testing.assert_array_equal(depth, ref_dep)


def test_bg_painted_in_place_channels_and_dtype_errors():
    v, t, c = synth.uv_grid_mesh(grid=5, res=48, seed=1)
    c4 = np.concatenate([c, c[:, :1]], 1)
    bg = np.full((48, 48, 4), 0.25, np.float32)
    out = f3d_render.render_colors(v * 0.5, t, c4, 48, 48, c=4, BG=bg)
    assert out is bg
    ref, _ = _cpu(v * 0.5, t, c4, 48, 48, 4, BG=np.full((48, 48, 4), 0.25, np.float32))
    np.testing.assert_array_equal(out, ref)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        mcc.render_colors_core(np.zeros((4, 4, 3), np.float64), np.zeros((3, 3), np.float32), np.zeros((1, 3), np.int32),
                               np.zeros((3, 3), np.float32), np.zeros((4, 4), np.float32), 3, 1, 4, 4, 3)


def test_full_size_properties_8k():
    """BASELINE config 4 at full size (8192^2, 120k triangles) through size-independent properties:
    idempotence (second pass over the same depth buffer changes nothing), coverage = union of the
    z=0 painter rule (every covered pixel's winner is the lowest triangle index -> equals CPU on a
    sampled band), and the uint8 epilogue."""
    res = 8192
    v, t, c = synth.uv_grid_mesh(grid=245, res=res, seed=0)
    dev = torch.device("cuda:0")
    d_v = torch.tensor(v, dtype=torch.float32, device=dev); d_t = torch.tensor(t, dtype=torch.int32, device=dev)
    d_c = torch.tensor(c, dtype=torch.float32, device=dev)
    img = torch.zeros((res, res, 3), device=dev); dep = torch.full((res, res), -999999.0, device=dev)
    ws = mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3)
    img2 = img.clone(); dep2 = dep.clone()
    mcc.render_colors_device(img2, d_v, d_t, d_c, dep2, res, res, 3, ws)
    assert torch.equal(img, img2) and torch.equal(dep, dep2)        # strict '>' makes a re-run a no-op
    # CPU check on two horizontal bands (incl. the top border rows): the reference painter at full size but
    # only over the triangles whose bbox touches the band -- every triangle that can cover a band pixel is
    # present and index order is preserved, so the band must match bit for bit (no translation: fp32
    # barycentrics are not translation invariant).
    ys = v[:, 1][t]
    for band in ((0, 40), (4000, 4064)):
        sel = (ys.max(1) >= band[0] - 1) & (ys.min(1) <= band[1] + 1)
        ref, _ = _cpu(v, t[sel], c, res, res, 3)
        got = img[band[0]:band[1]].cpu().numpy()
        np.testing.assert_array_equal(got, ref[band[0]:band[1]])
        del ref
    u8 = mcc.image_to_u8_device(img)
    assert torch.equal(u8, (img * 255).to(torch.uint8))
    # fused bake epilogue == reference's (render * 255).astype(np.uint8) (helpers.py:959) on a small mesh
    vs, ts, cs = synth.uv_grid_mesh(grid=20, res=256, seed=2)
    ref_small, _ = _cpu(vs, ts, cs, 256, 256, 3)
    np.testing.assert_array_equal(f3d_render.render_colors_u8(vs, ts, cs, 256, 256, 3), (ref_small * 255).astype(np.uint8))
    assert float((img.sum(-1) != 0).float().mean()) > 0.95


@pytest.mark.parametrize("grid,res", [(40, 300), (12, 256), (9, 514)])
@pytest.mark.parametrize("zmode", ["flat", "random"])
def test_fused_bake_path_float_and_u8(zmode, grid, res):
    """render_colors without BG takes the fused bake (f3d_bake_colors: fresh image, private constant depth, optional uint8):
    bit-identical to the reference followed by (image * 255).astype(uint8), for all-zero z (one key stage) and general z (two)."""
    v, t, c = synth.uv_grid_mesh(grid=grid, res=res, seed=7)
    if zmode == "random":
        v[:, 2] = np.random.default_rng(3).normal(size=v.shape[0])
    ref, _ = _cpu(v, t, c, res, res, 3)
    np.testing.assert_array_equal(f3d_render.render_colors(v, t, c, res, res, 3), ref)
    np.testing.assert_array_equal(f3d_render.render_colors_u8(v, t, c, res, res, 3), (ref * 255).astype(np.uint8))
    # an empty mesh still yields the zero image (render.py:66)
    assert not f3d_render.render_colors(v, t[:0], c, 64, 48, 3).any()


def test_row_bands_do_not_change_results():
    """Images beyond the scratch budget are resolved in bands of rows; force 16-row bands on a small image."""
    from topo4d_b200 import _lib
    v, t, c = synth.uv_grid_mesh(grid=23, res=200, seed=11)
    v[:, 2] = np.random.default_rng(5).normal(size=v.shape[0])
    ref, _ = _cpu(v, t, c, 200, 200, 3)
    try:
        _lib.lib().f3d_set_band_bytes(16 * 200 * 4)
        np.testing.assert_array_equal(f3d_render.render_colors(v, t, c, 200, 200, 3), ref)
        bg = np.full((200, 200, 3), 0.5, np.float32)
        ref_bg, _ = _cpu(v, t, c, 200, 200, 3, BG=bg.copy())
        np.testing.assert_array_equal(f3d_render.render_colors(v, t, c, 200, 200, 3, BG=bg), ref_bg)
    finally:
        _lib.lib().f3d_set_band_bytes(0)
