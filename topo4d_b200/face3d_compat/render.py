"""`render_colors` with the signature and behaviour of face3d/mesh/render.py:52-86, plus the fused uint8 bake
(`render_colors_u8`) that write_texture (helpers.py:953-960) can call instead of render + `*255` + astype."""
from __future__ import annotations

import numpy as np
import torch

from . import mesh_core_cython


def _device_render(vertices, triangles, colors, h, w, c, BG, device):
    """Uploads the mesh, initialises image / depth ON THE DEVICE (zeros or BG, -999999) and rasterizes.
    The reference allocates and casts on the host (render.py:66-77); values are identical."""
    dev = torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("topo4d_b200.face3d_compat: CUDA device required (there is no CPU path)")
    d_v = torch.from_numpy(np.ascontiguousarray(vertices, dtype=np.float32)).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(triangles, dtype=np.int32)).to(dev)
    d_c = torch.from_numpy(np.ascontiguousarray(colors, dtype=np.float32)).to(dev)
    if BG is None:
        d_img = torch.zeros((h, w, c), dtype=torch.float32, device=dev)
    else:
        d_img = torch.from_numpy(BG).to(dev)
    d_dep = torch.full((h, w), -999999.0, dtype=torch.float32, device=dev)
    mesh_core_cython.render_colors_device(d_img, d_v, d_t, d_c, d_dep, h, w, c)
    return d_img


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
        image = np.empty((h, w, c), dtype=np.float32)
    else:
        assert BG.shape[0] == h and BG.shape[1] == w and BG.shape[2] == c
        if BG.dtype != np.float32 or not BG.flags.c_contiguous:
            raise ValueError("Buffer dtype mismatch, expected 'float32' C-contiguous BG")   # as the typed Cython arg would
        image = BG
    d_img = _device_render(vertices, triangles, colors, h, w, c, BG, device)
    torch.from_numpy(image).copy_(d_img)                 # one D2H straight into the array that is returned
    return image


def render_colors_u8(vertices, triangles, colors, h, w, c=3, device="cuda"):
    """The whole write_texture body up to imsave (helpers.py:956-959): render, `*255`, astype(uint8) -- on the
    device, so only h*w*c bytes cross PCIe instead of 4x that."""
    d_img = _device_render(vertices, triangles, colors, h, w, c, None, device)
    out = np.empty((h, w, c), dtype=np.uint8)
    torch.from_numpy(out).copy_(mesh_core_cython.image_to_u8_device(d_img))
    return out
