"""`render_colors_core` with the exact calling convention of the reference's Cython shim
(face3d/mesh/cython/mesh_core_cython.pyx:64-77): C-contiguous float32/int32 NumPy buffers,
`image` [h,w,c] and `depth_buffer` [h,w] mutated IN PLACE, returns None.  The work runs on the GPU
(f3d_render_colors in libtopo4d_b200.so); host<->device copies happen here because the reference
caller (face3d/mesh/render.py:52-86 <- helpers.py:956) holds NumPy arrays.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib


def _check(name, a, dtype, ndim):
    # same failure mode as the typed Cython signature: "Buffer dtype mismatch" -> ValueError
    if not isinstance(a, np.ndarray):
        raise TypeError(f"Argument '{name}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
    if a.dtype != dtype:
        raise ValueError(f"Buffer dtype mismatch, expected '{np.dtype(dtype).name}' but got '{a.dtype.name}'")
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")


def render_colors_device(image, vertices, triangles, colors, depth_buffer, h, w, c, workspace=None):
    """Device-tensor path (no copies): all arguments are CUDA tensors; image/depth updated in place."""
    L = _lib.lib()
    dev = image.device
    ntri, nver = int(triangles.shape[0]), int(vertices.shape[0])
    need = L.f3d_workspace_bytes(ntri, h, w)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = L.f3d_render_colors(C.c_void_p(image.data_ptr()), C.c_void_p(vertices.data_ptr()),
                                   C.c_void_p(triangles.data_ptr()), C.c_void_p(colors.data_ptr()),
                                   C.c_void_p(depth_buffer.data_ptr()), nver, ntri, int(h), int(w), int(c),
                                   C.c_void_p(workspace.data_ptr()), workspace.numel(),
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if code != 0:
        raise RuntimeError(f"f3d_render_colors failed with code {code}")
    return workspace


def bake_colors_device(vertices, triangles, colors, h, w, c, u8=False, depth_init=-999999.0, workspace=None):
    """face3d/mesh/render.py:52-86 as one device pass: fresh image (every pixel written: winner colour or 0), private constant
    depth.  Returns a [h,w,c] float32 (or uint8 = (value*255) truncated, helpers.py:959) CUDA tensor."""
    L = _lib.lib()
    dev = vertices.device
    ntri, nver = int(triangles.shape[0]), int(vertices.shape[0])
    need = L.f3d_workspace_bytes(ntri, h, w)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    out = torch.empty((h, w, c), dtype=torch.uint8 if u8 else torch.float32, device=dev)
    with torch.cuda.device(dev):
        code = L.f3d_bake_colors(None if u8 else C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr()) if u8 else None,
                                 C.c_void_p(vertices.data_ptr()), C.c_void_p(triangles.data_ptr()), C.c_void_p(colors.data_ptr()),
                                 float(depth_init), nver, ntri, int(h), int(w), int(c), C.c_void_p(workspace.data_ptr()),
                                 workspace.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if code != 0:
        raise RuntimeError(f"f3d_bake_colors failed with code {code}")
    return out


def image_to_u8_device(image: torch.Tensor) -> torch.Tensor:
    """(image*255).astype(uint8) of helpers.py:959 on the device."""
    out = torch.empty(image.shape, dtype=torch.uint8, device=image.device)
    with torch.cuda.device(image.device):
        code = _lib.lib().f3d_image_to_u8(C.c_void_p(image.data_ptr()), C.c_void_p(out.data_ptr()), image.numel(),
                                          C.c_void_p(torch.cuda.current_stream(image.device).cuda_stream))
    if code != 0:
        raise RuntimeError(f"f3d_image_to_u8 failed with code {code}")
    return out


def render_colors_core(image, vertices, triangles, colors, depth_buffer, nver, ntri, h, w, c, device="cuda"):
    _check("image", image, np.float32, 3)
    _check("vertices", vertices, np.float32, 2)
    _check("triangles", triangles, np.int32, 2)
    _check("colors", colors, np.float32, 2)
    _check("depth_buffer", depth_buffer, np.float32, 2)
    if not torch.cuda.is_available():
        raise RuntimeError("topo4d_b200.face3d_compat: CUDA device required (there is no CPU path)")
    # the C-ABI host-pointer entry does the copies, the kernels and the in-place update (f3d_render_colors_host)
    with torch.cuda.device(torch.device(device)):
        code = _lib.lib().f3d_render_colors_host(C.c_void_p(image.ctypes.data), C.c_void_p(vertices.ctypes.data),
                                                 C.c_void_p(triangles.ctypes.data), C.c_void_p(colors.ctypes.data),
                                                 C.c_void_p(depth_buffer.ctypes.data), int(nver), int(ntri), int(h), int(w), int(c))
    if code != 0:
        raise RuntimeError(f"f3d_render_colors_host failed with code {code}")
    return None
