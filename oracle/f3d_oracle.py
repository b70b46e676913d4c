"""ctypes front-ends of the face3d oracles (TEST INFRASTRUCTURE -- see oracle/__init__.py).

  render_colors_port : oracle/f3d_oracle.c, our restatement of mesh_core.cpp:169-234
  render_colors_ref  : the reference's own C++ (oracle/_ref/libf3d_ref.so, built by oracle/Makefile
                       from /root/reference/face3d/mesh/cython/mesh_core.cpp where it lies)
Both mirror face3d/mesh/render.py:52-86: float32/int32 casts, image zeros (or BG, painted in
place), depth buffer initialised to -999999; they return (image, depth_buffer).
"""
import ctypes as C
import os

import numpy as np

from . import _HERE, lib_path

_port = None
_ref = None


def have_ref() -> bool:
    if os.path.exists(os.path.join(_HERE, "_ref", "libf3d_ref.so")):
        return True
    try:
        lib_path("libf3d_ref.so")
        return True
    except Exception:
        return False


def _sig(fn):
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int] * 5
    return fn


def _run(fn, vertices, triangles, colors, h, w, c, BG, depth_init):
    image = np.zeros((h, w, c), np.float32) if BG is None else BG
    assert image.dtype == np.float32 and image.flags.c_contiguous and image.shape == (h, w, c)
    depth = (np.zeros((h, w), np.float32) - 999999.0) if depth_init is None else depth_init
    v = np.ascontiguousarray(np.asarray(vertices).astype(np.float32))
    t = np.ascontiguousarray(np.asarray(triangles).astype(np.int32))
    col = np.ascontiguousarray(np.asarray(colors).astype(np.float32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn(p(image), p(v), p(t), p(col), p(depth), v.shape[0], t.shape[0], h, w, c)
    return image, depth


def render_colors_port(vertices, triangles, colors, h, w, c=3, BG=None, depth_init=None):
    global _port
    if _port is None:
        _port = _sig(C.CDLL(lib_path("libf3d_oracle.so")).f3d_port_render_colors)
    return _run(_port, vertices, triangles, colors, h, w, c, BG, depth_init)


def render_colors_ref(vertices, triangles, colors, h, w, c=3, BG=None, depth_init=None):
    global _ref
    if _ref is None:
        _ref = _sig(C.CDLL(lib_path("libf3d_ref.so")).f3d_ref_render_colors)
    return _run(_ref, vertices, triangles, colors, h, w, c, BG, depth_init)
