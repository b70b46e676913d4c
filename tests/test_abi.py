"""CPU: the C-ABI library builds/loads and exports every symbol include/topo4d_b200.h declares;
host-side argument validation and workspace arithmetic (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from topo4d_b200 import _lib, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "topo4d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gs|f3d|t4d)_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    names = _declared_functions()
    assert len(names) >= 12
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/topo4d_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in topo4d_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(names)


def test_struct_layouts_match_header_sizes():
    # natural alignment on LP64: GsProblem = 8 ints/floats (32) + i64 + 8 pointers + pointer + size_t
    assert C.sizeof(_lib.GsProblem) == 40 + 8 + 8 * 8 + 8 + 8
    assert C.sizeof(_lib.GsForwardOut) == 4 * 8
    assert C.sizeof(_lib.GsBackwardIO) == 12 * 8
    assert C.sizeof(_lib.GsStatus) == 32
    assert C.sizeof(_lib.GsWorkspaceView) == 7 * 8 + 8


def test_workspace_bytes_arithmetic():
    L = _lib.lib()
    a = L.gs_workspace_bytes(60000, 24, 1080, 1920, 4_000_000)
    b = L.gs_workspace_bytes(60000, 24, 1080, 1920, 8_000_000)
    assert b - a >= 4_000_000 * 60 and b - a < 4_000_000 * 61        # 8 pair + 4 id + 48 record per instance
    c = L.gs_workspace_bytes(60000, 1, 1080, 1920, 4_000_000)
    per_view = (a - c) / 23
    assert per_view >= 1080 * 1920 * 8 + 60000 * 97                  # final_T + n_contrib + geom + grad2d + clamp
    assert L.gs_workspace_bytes(-1, 1, 8, 8, 0) == 0 and L.gs_workspace_bytes(1, 0, 8, 8, 0) == 0
    # face3d: header + two u32 key planes for one band of rows (<= 1 GiB per plane: the whole image at 8192^2)
    # ... plus one 112-byte shade record per triangle (up to 2 M triangles)
    assert L.f3d_workspace_bytes(10, 8192, 8192) == 256 + 2 * 8192 * 8192 * 4 + 10 * 112
    assert L.f3d_workspace_bytes(3_000_000, 32768, 32768) == 256 + 2 * 8192 * 32768 * 4
    assert L.f3d_workspace_bytes(0, 64, 48) == 256 + 2 * 64 * 48 * 4 and L.f3d_workspace_bytes(10, 0, 8) == 0


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.gs_forward(None, None, None) == _lib.GS_E_BAD_ARGS
    pr = _lib.GsProblem(N=10, V=1, H=16, W=16, cap_instances=100)
    out = _lib.GsForwardOut()
    assert L.gs_forward(C.byref(pr), C.byref(out), None) == _lib.GS_E_BAD_ARGS           # no workspace / cameras
    assert L.gs_backward(C.byref(pr), None, None) == _lib.GS_E_BAD_ARGS
    assert L.f3d_render_colors(None, None, None, None, None, 0, 0, 8, 8, 3, None, 0, None) == -1
    assert L.f3d_bake_colors(None, None, None, None, None, 0.0, 0, 0, 8, 8, 3, None, 0, None) == -1
    assert L.f3d_render_colors_host(None, None, None, None, None, 0, 0, 8, 8, 3) == -1
    assert b"capacity" in L.gs_last_error(_lib.GS_E_OVERFLOW)
    with pytest.raises(_lib.GsError):
        _lib.check(_lib.GS_E_WORKSPACE_SMALL, "x")


def test_python_surface_errors_and_camera_packing():
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    assert GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg",
                                                     "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos",
                                                     "prefiltered", "debug")                 # helpers.py:73-86 order
    r = GaussianRasterizer(raster_settings=None)
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z, means2D=z, opacities=z[:, :1], shs=z, colors_precomp=z, scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="precomputed 3D covariance"):
        r(means3D=z, means2D=z, opacities=z[:, :1], colors_precomp=z)
    cams = synth.ring_cameras(4, w=64, h=48)
    p = engine.pack_cameras_numpy(cams, (0.1, 0.2, 0.3))
    assert p.shape == (4, 48) and p.dtype == np.float32
    np.testing.assert_array_equal(p[2, 0:16], cams[2].viewmatrix.reshape(16))
    np.testing.assert_allclose(p[1, 35:38], [0.1, 0.2, 0.3])
    assert p[0, 38] == np.float32(cams[0].tanfovx) and (p[:, 40:] == 0).all()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        engine.forward(z, z[:, :1], torch.zeros(1, 48), 16, 16, colors_precomp=z, scales=z, rotations=torch.zeros(4, 4))


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under topo4d_b200/ or diff_gaussian_rasterization/ may touch it."""
    bad = []
    for pkg in ("topo4d_b200", "diff_gaussian_rasterization"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|oracle/_|libgs_oracle|libf3d_(ref|oracle)", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_loop_tail_abi_validation_without_gpu():
    """t4d_* entry points: struct layouts, workspace arithmetic and argument validation (no compute)."""
    import torch
    from topo4d_b200 import dense, losses, optim
    L = _lib.lib()
    # LP64 layouts of the header structs: 3 int32 + 2 float (20, padded to 24) + 8 pointers + pointer + size_t
    assert C.sizeof(_lib.T4dImageLoss) == 24 + 8 * 8 + 8 + 8
    # 6 pointers + int64 + 2 int32 + float (+4 pad) + 2 pointers
    assert C.sizeof(_lib.T4dAdamSegment) == 6 * 8 + 8 + 4 + 4 + 8 + 2 * 8
    nblk = ((1920 + 31) // 32) * ((1080 + 31) // 32)
    ws = L.t4d_image_loss_workspace_bytes(2, 1080, 1920)
    assert ws >= 2 * 9 * 1080 * 1920 * 4 + 2 * 2 * 3 * nblk * 8 and ws < 2 * 9 * 1080 * 1920 * 4 + 2 * 2 * 3 * nblk * 8 + 3 * 256
    assert L.t4d_image_loss_workspace_bytes(0, 8, 8) == 0
    assert L.t4d_image_loss(None, None) == _lib.GS_E_BAD_ARGS
    bad = _lib.T4dImageLoss(V=1, H=8, W=8, w_l1=0.8, w_ssim=0.2)                      # no pointers
    assert L.t4d_image_loss(C.byref(bad), None) == _lib.GS_E_BAD_ARGS
    assert L.t4d_adam_step(None, 1, 0.9, 0.999, 1e-15, None) == _lib.GS_E_BAD_ARGS
    seg = (_lib.T4dAdamSegment * 1)(_lib.T4dAdamSegment(count=16, row_width=1, step=0, lr=0.1))      # step 0 / NULL tensors
    assert L.t4d_adam_step(seg, 1, 0.9, 0.999, 1e-15, None) == _lib.GS_E_BAD_ARGS
    assert L.t4d_adam_step(seg, _lib.T4D_ADAM_MAX_SEGMENTS + 1, 0.9, 0.999, 1e-15, None) == _lib.GS_E_BAD_ARGS
    assert L.t4d_dense_attribute(None, 4, 3, None, None, None, 2, None, None) == _lib.GS_E_BAD_ARGS
    assert L.t4d_dense_attribute(None, 0, 3, None, None, None, 0, None, None) == 0      # nothing to do
    # the host mirrors refuse CPU tensors: there is no CPU fallback on the product path
    x = torch.rand(3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        losses.image_loss(x, x)
    with pytest.raises(ValueError):
        losses.image_loss(x, x, torch.zeros(3), None)
    with pytest.raises(NotImplementedError):
        losses.calc_ssim(x, x, window_size=7)
    p = torch.nn.Parameter(torch.zeros(4, 3))
    p.grad = torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        optim.FusedAdam([{"params": [p], "name": "means3D", "lr": 0.1}], lr=0.0, eps=1e-15).step()
    with pytest.raises(ValueError):
        optim.FusedAdam([{"params": [p], "lr": 0.1, "eps": 1e-8}, {"params": [torch.nn.Parameter(torch.zeros(2))], "lr": 0.1, "eps": 1e-3}])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dense.compute_vertex_attribute_by_weight_2({}, torch.zeros(4, 3))


def test_debug_mode_leaves_a_snapshot_when_the_forward_fails(tmp_path, monkeypatch):
    """settings.debug=True: like upstream's wrapper, a failing forward writes its inputs to snapshot_fw.dump before the
    exception propagates (here the failure is the CUDA-only check on CPU tensors); debug=False writes nothing."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    monkeypatch.chdir(tmp_path)
    cam = synth.front_camera(32, 24)

    def settings(debug):
        return GaussianRasterizationSettings(image_height=24, image_width=32, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                             bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=torch.tensor(cam.viewmatrix).reshape(1, 4, 4),
                                             projmatrix=torch.tensor(cam.projmatrix).reshape(1, 4, 4), sh_degree=0,
                                             campos=torch.tensor(cam.campos), prefiltered=False, debug=debug)
    z = torch.rand(5, 3)
    kw = dict(means3D=z, means2D=torch.zeros(5, 3), opacities=torch.ones(5, 1), colors_precomp=z, scales=z, rotations=torch.rand(5, 4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        GaussianRasterizer(raster_settings=settings(False))(**kw)
    assert not os.path.exists("snapshot_fw.dump")
    with pytest.raises(RuntimeError, match="CUDA-only"):
        GaussianRasterizer(raster_settings=settings(True))(**kw)
    snap = torch.load("snapshot_fw.dump")
    assert torch.equal(snap["means3D"], z) and snap["image_height"] == 24 and snap["sh"] is None
