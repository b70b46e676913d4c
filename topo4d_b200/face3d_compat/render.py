"""`render_colors` with the signature and behaviour of face3d/mesh/render.py:52-86."""
from __future__ import annotations

import numpy as np

from . import mesh_core_cython


def render_colors(vertices, triangles, colors, h, w, c=3, BG=None):
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
        image = np.zeros((h, w, c), dtype=np.float32)
    else:
        assert BG.shape[0] == h and BG.shape[1] == w and BG.shape[2] == c
        image = BG
    depth_buffer = np.zeros([h, w], dtype=np.float32, order='C') - 999999.
    vertices = vertices.astype(np.float32).copy()
    triangles = triangles.astype(np.int32).copy()
    colors = colors.astype(np.float32).copy()
    mesh_core_cython.render_colors_core(image, vertices, triangles, colors, depth_buffer,
                                        vertices.shape[0], triangles.shape[0], h, w, c)
    return image
