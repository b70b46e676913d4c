"""Upstream-style CUDA arm (benchmark denominator; see us_raster.cu).  One view per call, like the reference renders
(train.py:307): K1 by this repository's preprocess kernel, K2-K7 by the upstream-style pipeline in libupstream_style.so
(CUB scan + host read of num_rendered, duplicateWithKeys, 64-bit CUB radix sort, tile ranges, 16x16-block blend, per-thread
global atomics in the backward), K8/K9 by this repository's preprocess backward.  Test / tools code only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

from topo4d_b200 import _lib, engine
from topo4d_b200._lib import GsBackwardIO, GsForwardOut, GsProblem, GsWorkspaceView

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libupstream_style.so")
_H = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "us_raster.cu")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                        "-shared", "-o", LIB, src], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB


def lib():
    global _H
    if _H is None:
        h = C.CDLL(LIB if os.path.exists(LIB) else build())
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        h.us_forward.restype = i
        h.us_forward.argtypes = [vp, vp, i, i, i, f, f, f, vp, vp, vp, vp]
        h.us_backward.restype = i
        h.us_backward.argtypes = [vp, i, i, i, f, f, f, vp, vp, vp, vp, vp]
        h.us_num_rendered.restype = C.c_longlong
        _H = h
    return _H


class View:
    """State of one rendered view (what upstream keeps in geomBuffer / binningBuffer / imgBuffer)."""
    def __init__(self, problem, ws, geom_ptr, grad2d_ptr, radii, keep):
        self.problem, self.ws, self.geom_ptr, self.grad2d_ptr, self.radii, self.keep = problem, ws, geom_ptr, grad2d_ptr, radii, keep


def forward(t: dict, cam: torch.Tensor, H: int, W: int, sh_degree: int, bg=(0.0, 0.0, 0.0)):
    """t: dict of CUDA tensors (means3D, opacities, scales, rotations, shs | colors_precomp); cam: [1,48]."""
    L, U = _lib.lib(), lib()
    dev = t["means3D"].device
    N = int(t["means3D"].shape[0])
    shs, col = t.get("shs"), t.get("colors_precomp")
    M = 0 if shs is None else int(shs.shape[1])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nbytes = L.gs_workspace_bytes(N, 1, H, W, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    p = engine._ptr
    pr = GsProblem(N, 1, H, W, int(sh_degree), M, 1.0, 0, 0, 0, 0, p(t["means3D"]), p(shs), p(col), p(t["opacities"]), p(t["scales"]),
                   p(t["rotations"]), None, p(cam), p(ws), nbytes)
    imgs = torch.empty(5 * H * W, dtype=torch.float32, device=dev)
    color, depth, alpha = imgs[:3 * H * W].view(3, H, W), imgs[3 * H * W:4 * H * W].view(1, H, W), imgs[4 * H * W:].view(1, H, W)
    radii = torch.empty(N, dtype=torch.int32, device=dev)
    out = GsForwardOut(p(color), p(depth), p(alpha), p(radii))
    _lib.check(L.gs_forward_stages(C.byref(pr), C.byref(out), _lib.GS_FWD_PREPROCESS, stream), "preprocess")
    wv = GsWorkspaceView()
    _lib.check(L.gs_workspace_view(C.byref(pr), C.byref(wv)), "view")
    rc = U.us_forward(wv.geom_records, p(radii), N, H, W, float(bg[0]), float(bg[1]), float(bg[2]), p(color), p(depth), p(alpha), stream)
    assert rc == 0, rc
    return color, radii, depth, alpha, View(pr, ws, wv.geom_records, wv.grad2d, radii, (t, cam))


def backward(v: View, g_color, g_depth, g_alpha, flat: torch.Tensor, bg=(0.0, 0.0, 0.0)):
    """Gradients of one view into `flat` (engine.flat_layout order); returns the named views of it."""
    L, U = _lib.lib(), lib()
    t, cam = v.keep
    dev = t["means3D"].device
    pr = v.problem
    N, H, W, M = pr.N, pr.H, pr.W, pr.sh_coeffs
    use_sh = t.get("shs") is not None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = engine._ptr
    rc = U.us_backward(v.geom_ptr, N, H, W, float(bg[0]), float(bg[1]), float(bg[2]), p(g_color), p(g_depth), p(g_alpha), v.grad2d_ptr, stream)
    assert rc == 0, rc
    seg = engine.flat_views(flat, N, M, use_sh, False)
    io = GsBackwardIO(p(g_color), p(g_depth), p(g_alpha), p(v.radii), p(seg["means3D"]), p(seg["means2D"]),
                      p(seg["shs"]) if use_sh else None, None if use_sh else p(seg["colors_precomp"]), p(seg["opacities"]),
                      p(seg["scales"]), p(seg["rotations"]), None)
    _lib.check(L.gs_backward_stages(C.byref(pr), C.byref(io), _lib.GS_BWD_PREPROCESS, stream), "preprocess_bwd")
    return seg


def num_rendered() -> int:
    return int(lib().us_num_rendered())
