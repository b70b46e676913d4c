"""Bake output path and one-off densification cache (the parts around the kernels).

Reference, once per frame (train.py:755 -> helpers.py:965-998):
  ``dense_colors = params["dense_rgb_colors"].clamp(0, 1).cpu().numpy()``                    D2H of every dense colour
  ``duplicate_texture_vertex_color_2(variables, dense_colors[:N])`` (helpers.py:930-941)     a Python dict over every UV of every
                                                                                             vertex, rebuilt on EVERY call, then a
                                                                                             list comprehension gather
  ``write_texture`` (helpers.py:953-960): ``process_uv``, ``render_colors``, ``*255``, ``astype(uint8)``, ``io.imsave`` (PNG encode
  on the critical path of the frame loop).

Here the UV -> vertex map is built once and cached on ``variables`` (it depends only on the topology), the gather and the clamp
run on the device, the bake is the fused ``f3d_bake_colors`` with uint8 output, and the PNG is encoded on a background thread
so the next frame's optimisation starts immediately (``TextureWriter.wait()`` before exit).  Same names and argument meaning as
the reference functions, so the switch is an import.

``cached_call`` is the "save the calculation result, and directly load it" the reference's own comment asks for
(train.py:226-229) around ``build_dense_vertices_2`` (helpers.py:602-654: minutes of Python loops whose result depends only on
topology and density): a content-addressed ``.npz`` next to the data.
"""
from __future__ import annotations

import hashlib
import os
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Sequence

import numpy as np
import torch

_KEY = "_t4d_uv_vertex_index"


def build_uv_vertex_index(variables) -> np.ndarray:
    """The index map duplicate_texture_vertex_color_2 derives on every call: for each entry of ``uvs_ori`` the vertex whose
    ``uvs_texture_ori`` list contains that UV (helpers.py:930-941; later vertices overwrite earlier ones, as the dict does)."""
    uv_dict = {}
    for idx, uvs_ in enumerate(variables["uvs_texture_ori"]):
        for uv in uvs_:
            uv_dict[tuple(uv)] = idx
    return np.asarray([uv_dict[tuple(uv)] for uv in variables["uvs_ori"]], dtype=np.int64)


def duplicate_texture_vertex_color_2(variables, colors):
    """Drop-in for helpers.py:930-941 (vertices on the UV seam own several UV coordinates: one colour row per UV).  The index
    map is computed once per ``variables``; `colors` may be a NumPy array (returns an array, like ``np.array(reference list)``)
    or a torch tensor on any device (returns a tensor there: the gather stays on the GPU)."""
    idx = variables.get(_KEY)
    if idx is None:
        idx = variables[_KEY] = build_uv_vertex_index(variables)
    if torch.is_tensor(colors):
        key = (_KEY, str(colors.device))
        dev_idx = variables.get(key)
        if dev_idx is None:
            dev_idx = variables[key] = torch.from_numpy(idx).to(colors.device)
        return colors.index_select(0, dev_idx)
    return np.asarray(colors)[idx]


def process_uv(uv_coords, uv_h: int = 256, uv_w: int = 256):
    """helpers.py:945-950 without mutating the caller's array (the reference scales `uv_coords` in place, which is why its
    caller passes a copy): pixel coordinates with the v axis flipped, z = 0."""
    uv = np.asarray(uv_coords, dtype=np.float64)
    out = np.zeros((uv.shape[0], 3), np.float64)
    out[:, 0] = uv[:, 0] * (uv_w - 1)
    out[:, 1] = uv_h - uv[:, 1] * (uv_h - 1) - 1
    return out


class TextureWriter:
    """``write_texture`` with the PNG encode off the critical path.  One instance per run; ``wait()`` joins the pending files."""

    def __init__(self, workers: int = 2):
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="t4d-texture")
        self._pending: list[Future] = []
        self._lock = threading.Lock()

    @staticmethod
    def _save(path: str, image_u8: np.ndarray) -> str:
        from PIL import Image
        Image.fromarray(np.squeeze(image_u8)).save(path)
        return path

    def write_texture(self, path, uvs, colors, faces, res: int = 1024, device="cuda") -> Future:
        """helpers.py:953-960: bake the per-vertex colours into a res x res uint8 texture and save it as `path`.  `colors` may be
        a device tensor (no D2H of the colours); triangles only (the reference triangulates before, helpers.py:657-667).
        Returns a Future of the written path; the bake itself has finished when this returns."""
        from .face3d_compat import mesh_core_cython as mcc
        from .face3d_compat import render as f3d
        dev = torch.device(device)
        uv = process_uv(uvs, res, res)
        if torch.is_tensor(colors):
            _, d_v, d_t, _ = f3d._upload(uv, np.asarray(faces), np.zeros((1, 3), np.float32), dev)
            d_c = colors.to(device=dev, dtype=torch.float32).contiguous()
            tex = f3d._to_host(mcc.bake_colors_device(d_v, d_t, d_c, res, res, 3, u8=True))
        else:
            tex = f3d.render_colors_u8(uv, np.asarray(faces), np.asarray(colors), res, res, 3, device=device)
        fut = self._pool.submit(self._save, path, tex)
        with self._lock:
            self._pending = [f for f in self._pending if not f.done()] + [fut]
        return fut

    def wait(self) -> None:
        with self._lock:
            pending, self._pending = self._pending, []
        for f in pending:
            f.result()


def dense_texture_colors(params, variables) -> torch.Tensor:
    """The colour table save_mesh assembles before write_texture (helpers.py:991-996), on the device:
    clamp(dense_rgb_colors, 0, 1), seam duplication of the first N rows, the densified rows appended."""
    dense = params["dense_rgb_colors"].detach().clamp(0.0, 1.0)
    n = int(params["means3D"].shape[0])
    return torch.cat([duplicate_texture_vertex_color_2(variables, dense[:n]), dense[n:]], dim=0)


def cached_call(fn: Callable, arrays: Sequence, cache_dir: str, tag: str, extra=()):
    """Content-addressed disk cache for a pure function of arrays returning a tuple of arrays, e.g.

        out = cached_call(lambda: build_dense_vertices_2(variables, vertices, quad_faces, quad_faces_idx, dense_num, pt_uvs),
                          [vertices, quad_faces, quad_faces_idx, variables["uvs_ori"], ...], cache_dir, "dense", extra=(dense_num,))

    The key hashes the bytes, shapes and dtypes of `arrays` and the repr of `extra`; a hit loads the ``.npz`` instead of calling."""
    h = hashlib.sha256(repr(tuple(extra)).encode())
    for a in arrays:
        a = np.ascontiguousarray(np.asarray(a))
        h.update(str((a.shape, a.dtype.str)).encode())
        h.update(a.tobytes())
    path = os.path.join(cache_dir, f"{tag}_{h.hexdigest()[:20]}.npz")
    if os.path.exists(path):
        with np.load(path, allow_pickle=False) as z:
            return tuple(z[f"arr_{i}"] for i in range(len(z.files)))
    out = fn()
    out = out if isinstance(out, tuple) else (out,)
    os.makedirs(cache_dir, exist_ok=True)
    tmp = path + f".tmp{os.getpid()}.npz"
    np.savez(tmp, *[np.asarray(o) for o in out])
    os.replace(tmp, path)
    return tuple(np.asarray(o) for o in out)
