"""Bake output path helpers (topo4d_b200/texture.py) against golden vectors produced by executing the reference's own
duplicate_texture_vertex_color_2 / process_uv (tests/golden/make_golden_texture.py), plus the disk cache and the writer."""
import os

import numpy as np
import pytest
import torch

from topo4d_b200 import texture

G = os.path.join(os.path.dirname(__file__), "golden", "texture.npz")


def _variables(g):
    per, flat = g["per"], g["uvs_texture_flat"]
    uvs_texture, o = [], 0
    for k in per:
        uvs_texture.append([tuple(uv) for uv in flat[o:o + k]])
        o += k
    return {"uvs_ori": g["uvs_ori"], "uvs_texture_ori": uvs_texture}


def test_duplicate_texture_vertex_color_and_process_uv_match_reference():
    g = np.load(G)
    variables = _variables(g)
    out = texture.duplicate_texture_vertex_color_2(variables, g["colors"])
    np.testing.assert_array_equal(out, g["ref_colors"])
    assert texture._KEY in variables                                            # the map is built once ...
    out_t = texture.duplicate_texture_vertex_color_2(variables, torch.from_numpy(g["colors"]))
    np.testing.assert_array_equal(out_t.numpy(), g["ref_colors"])              # ... and serves tensors too
    uv = g["uv_in"].copy()
    np.testing.assert_array_equal(texture.process_uv(uv, 1024, 1024), g["ref_uv"])
    np.testing.assert_array_equal(uv, g["uv_in"])                               # unlike the reference, the input is not modified


def test_cached_call_hits_and_misses(tmp_path):
    calls = []

    def expensive():
        calls.append(1)
        return np.arange(6).reshape(2, 3), np.ones(4, np.float32)

    a, b = np.arange(5), np.eye(2)
    r1 = texture.cached_call(expensive, [a, b], str(tmp_path), "dense", extra=(30,))
    r2 = texture.cached_call(expensive, [a, b], str(tmp_path), "dense", extra=(30,))
    assert len(calls) == 1 and all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(r1, r2))
    texture.cached_call(expensive, [a, b], str(tmp_path), "dense", extra=(31,))                  # other density: recomputed
    texture.cached_call(expensive, [a + 1, b], str(tmp_path), "dense", extra=(30,))              # other topology: recomputed
    assert len(calls) == 3


@pytest.mark.gpu
def test_texture_writer_matches_reference_write_texture(tmp_path):
    """write_texture (helpers.py:953-960): process_uv -> render_colors -> *255 -> uint8 -> PNG, device colours, async save."""
    from PIL import Image
    from oracle import f3d_oracle
    from topo4d_b200 import synth
    v, t, c = synth.uv_grid_mesh(grid=24, res=256, seed=4, extras=False)
    uvs = np.stack([v[:, 0] / 255.0, 1.0 - v[:, 1] / 255.0], 1)                # inverse of process_uv at res 256
    cpu = (f3d_oracle.render_colors_ref if f3d_oracle.have_ref() else f3d_oracle.render_colors_port)
    ref, _ = cpu(texture.process_uv(uvs, 256, 256), t, c, 256, 256, 3)
    w = texture.TextureWriter()
    fut = w.write_texture(str(tmp_path / "face.png"), uvs, torch.tensor(c, dtype=torch.float32, device="cuda:0"), t, res=256)
    w.wait()
    assert fut.done()
    np.testing.assert_array_equal(np.array(Image.open(tmp_path / "face.png")), (ref * 255).astype(np.uint8))
