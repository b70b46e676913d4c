"""Host-side mirror of the reference's dense-mesh attribute interpolation over the C ABI.

``compute_vertex_attribute_by_weight_2(variables, attribute)`` keeps the reference's name and arguments
(helpers.py:237-253): ``variables`` holds the NumPy topology built once by initialize_dense_params
(train.py:238-243: 'dense_quad_faces' [F,4] int, 'dense_vertex_father' [M,1] int32, 'dense_vertex_weight' [M,4] float64,
'dense_vertex' [n_dense,3]).  The reference takes / returns NumPy float64 and its caller (train.py:504-506) wraps the
call in ``.cpu().numpy()`` / ``torch.from_numpy(...).cuda().float()``; here ``attribute`` is a CUDA tensor and the result
is the float32 CUDA tensor that whole expression produces, bit for bit, with no host round trip.  CUDA-only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_TOPO_CACHE: dict[tuple, tuple] = {}


def _device_topology(variables, dev):
    qf, vf, w = variables["dense_quad_faces"], variables["dense_vertex_father"], variables["dense_vertex_weight"]
    key = (id(qf), id(vf), id(w), dev.index)
    hit = _TOPO_CACHE.get(key)
    if hit is None:
        q = torch.from_numpy(np.ascontiguousarray(np.asarray(qf).reshape(-1, 4), dtype=np.int32)).to(dev)
        f = torch.from_numpy(np.ascontiguousarray(np.asarray(vf).reshape(-1), dtype=np.int32)).to(dev)
        ww = torch.from_numpy(np.ascontiguousarray(np.asarray(w).reshape(-1, 4), dtype=np.float64)).to(dev)
        if f.numel() != ww.shape[0]:
            raise ValueError("dense_vertex_father and dense_vertex_weight disagree on the number of new vertices")
        hit = _TOPO_CACHE[key] = (q, f, ww, (qf, vf, w))          # keep the sources alive: id() keys stay unique
        if len(_TOPO_CACHE) > 8:
            _TOPO_CACHE.pop(next(iter(_TOPO_CACHE)))
    return hit[:3]


def compute_vertex_attribute_by_weight_2(variables, attribute: torch.Tensor) -> torch.Tensor:
    if not isinstance(attribute, torch.Tensor) or not attribute.is_cuda:
        raise RuntimeError("topo4d_b200: compute_vertex_attribute_by_weight_2 takes a CUDA tensor; there is no CPU path")
    dev = attribute.device
    attr = attribute.detach()
    if attr.dtype is not torch.float32 or not attr.is_contiguous():
        attr = attr.float().contiguous()
    n_base, ch = int(attr.shape[0]), int(attr.shape[1])
    q, f, w = _device_topology(variables, dev)
    n_dense = int(variables["dense_vertex"].shape[0])
    n_new = n_dense - n_base
    if n_new != f.numel():
        raise ValueError(f"dense_vertex has {n_dense} rows but base ({n_base}) + interpolated ({f.numel()}) differ")
    out = torch.empty((n_dense, ch), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().t4d_dense_attribute(C.c_void_p(attr.data_ptr()), n_base, ch, C.c_void_p(q.data_ptr()),
                                                  C.c_void_p(f.data_ptr()), C.c_void_p(w.data_ptr()), n_new,
                                                  C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "t4d_dense_attribute")
    return out
