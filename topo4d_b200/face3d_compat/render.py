"""`render_colors` with the signature and behaviour of face3d/mesh/render.py:52-86, plus the fused uint8 bake
(`render_colors_u8`) that write_texture (helpers.py:953-960) can call instead of render + `*255` + astype."""
from __future__ import annotations

import numpy as np
import torch

from . import mesh_core_cython


def _upload(vertices, triangles, colors, device):
    dev = torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("topo4d_b200.face3d_compat: CUDA device required (there is no CPU path)")
    d_v = torch.from_numpy(np.ascontiguousarray(vertices, dtype=np.float32)).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(triangles, dtype=np.int32)).to(dev)
    d_c = torch.from_numpy(np.ascontiguousarray(colors, dtype=np.float32)).to(dev)
    return dev, d_v, d_t, d_c


def _device_render(vertices, triangles, colors, h, w, c, BG, device, u8=False):
    """Uploads the mesh and rasterizes ON THE DEVICE.  Without BG this is the fused bake (fresh image, private constant depth:
    nothing but the result is written); with BG the caller's image is painted into, with a -999999 depth plane like the
    reference's (render.py:66-77); values are identical."""
    dev, d_v, d_t, d_c = _upload(vertices, triangles, colors, device)
    if BG is None:
        return mesh_core_cython.bake_colors_device(d_v, d_t, d_c, h, w, c, u8=u8)
    d_img = torch.from_numpy(BG).to(dev)
    d_dep = torch.full((h, w), -999999.0, dtype=torch.float32, device=dev)
    mesh_core_cython.render_colors_device(d_img, d_v, d_t, d_c, d_dep, h, w, c)
    return mesh_core_cython.image_to_u8_device(d_img) if u8 else d_img


def _to_host(d_img: torch.Tensor) -> np.ndarray:
    """Device result -> NumPy array in page-locked memory from PyTorch's caching host allocator: the copy runs at PCIe speed
    and, from the second bake on, the 805 MB block is reused -- no page faults of a fresh np.empty, no pageable staging.  The
    block returns to the cache when the array is garbage-collected."""
    host = torch.empty(d_img.shape, dtype=d_img.dtype, pin_memory=True)
    host.copy_(d_img, non_blocking=True)
    torch.cuda.current_stream(d_img.device).synchronize()
    return host.numpy()


def render_colors(vertices, triangles, colors, h, w, c=3, BG=None, device="cuda"):
    ''' render mesh with colors
    Args:
        vertices: [nver, 3]   (x, y in pixel coordinates, z: larger = nearer)
        triangles: [ntri, 3]
        colors: [nver, 3]
        h: height
        w: width
        c: channel
        BG: background image (painted into, like the reference)
    Returns:
        image: [h, w, c] float32
    '''
    if BG is None:
        return _to_host(_device_render(vertices, triangles, colors, h, w, c, None, device))
    assert BG.shape[0] == h and BG.shape[1] == w and BG.shape[2] == c
    if BG.dtype != np.float32 or not BG.flags.c_contiguous:
        raise ValueError("Buffer dtype mismatch, expected 'float32' C-contiguous BG")   # as the typed Cython arg would
    d_img = _device_render(vertices, triangles, colors, h, w, c, BG, device)
    torch.from_numpy(BG).copy_(d_img)                    # painted INTO the caller's array, like the reference (render.py:70-71)
    return BG


def render_colors_u8(vertices, triangles, colors, h, w, c=3, device="cuda"):
    """The whole write_texture body up to imsave (helpers.py:956-959): render, `*255`, astype(uint8) -- on the
    device, so only h*w*c bytes cross PCIe instead of 4x that."""
    return _to_host(_device_render(vertices, triangles, colors, h, w, c, None, device, u8=True))
