"""ctypes front-end of oracle/gs_oracle.c (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Mirrors the op's call shape (reference helpers.py:73-112): one camera view per call,
numpy float32 in / out.  PARITY UNPINNED (see the header of gs_oracle.c).
"""
import ctypes as C
import numpy as np

from . import lib_path

_lib = None


def _L():
    global _lib
    if _lib is None:
        lib = C.CDLL(lib_path("libgs_oracle.so"))
        fp = C.c_void_p
        lib.gso_forward.restype = C.c_void_p
        lib.gso_forward.argtypes = [C.c_int] * 5 + [fp] * 7 + [C.c_float, fp, fp, fp, C.c_float, C.c_float, fp] + [fp] * 4
        lib.gso_free.argtypes = [C.c_void_p]
        lib.gso_num_rendered.restype = C.c_int64
        lib.gso_num_rendered.argtypes = [C.c_void_p]
        lib.gso_get_geometry.argtypes = [C.c_void_p] + [fp] * 8
        lib.gso_get_binning.argtypes = [C.c_void_p] + [fp] * 3
        lib.gso_get_image_state.argtypes = [C.c_void_p] + [fp] * 2
        lib.gso_backward.argtypes = [C.c_void_p] + [fp] * 12 + [C.c_int]
        lib.gso_mark_visible.argtypes = [C.c_int, fp, fp, fp]
        lib.gso_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(_L().gso_num_threads())


class GsOracleState:
    """Result of one forward; keeps the C state alive for backward / index queries."""

    def __init__(self, handle, N, M, H, W, use_sh, use_cov):
        self._h = handle
        self._free = _L().gso_free
        self.N, self.M, self.H, self.W = N, M, H, W
        self.use_sh, self.use_cov = use_sh, use_cov
        self.gx, self.gy = (W + 15) // 16, (H + 15) // 16

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_free", None):
            self._free(self._h)
            self._h = None

    @property
    def num_rendered(self) -> int:
        return int(_L().gso_num_rendered(self._h))

    def geometry(self):
        N = max(self.N, 1)
        out = dict(depth=np.zeros(N, np.float32), xy=np.zeros((N, 2), np.float32),
                   conic_opacity=np.zeros((N, 4), np.float32), rgb=np.zeros((N, 3), np.float32),
                   cov3d=np.zeros((N, 6), np.float32), rect=np.zeros((N, 4), np.int32),
                   tiles_touched=np.zeros(N, np.uint32), clamped=np.zeros((N, 3), np.uint8))
        _L().gso_get_geometry(self._h, *[_p(out[k]) for k in
                                         ("depth", "xy", "conic_opacity", "rgb", "cov3d", "rect", "tiles_touched", "clamped")])
        return {k: v[:self.N] for k, v in out.items()}

    def binning(self):
        I = max(self.num_rendered, 1)
        keys = np.zeros(I, np.uint64)
        vals = np.zeros(I, np.uint32)
        ranges = np.zeros((self.gx * self.gy, 2), np.uint32)
        _L().gso_get_binning(self._h, _p(keys), _p(vals), _p(ranges))
        n = self.num_rendered
        return dict(keys=keys[:n], ids=vals[:n], ranges=ranges)

    def image_state(self):
        final_T = np.zeros((self.H, self.W), np.float32)
        n_contrib = np.zeros((self.H, self.W), np.uint32)
        _L().gso_get_image_state(self._h, _p(final_T), _p(n_contrib))
        return dict(final_T=final_T, n_contrib=n_contrib)

    def backward(self, g_color, g_depth=None, g_alpha=None, return_acc2d=False, f32_replay=False):
        H, W, N, M = self.H, self.W, self.N, self.M
        g_color = _f32(g_color).reshape(3, H, W)
        g_depth = np.zeros((H, W), np.float32) if g_depth is None else _f32(g_depth).reshape(H, W)
        g_alpha = np.zeros((H, W), np.float32) if g_alpha is None else _f32(g_alpha).reshape(H, W)
        g = dict(means3D=np.zeros((N, 3), np.float32), means2D=np.zeros((N, 3), np.float32),
                 shs=np.zeros((N, max(M, 1), 3), np.float32), colors_precomp=np.zeros((N, 3), np.float32),
                 opacities=np.zeros((N, 1), np.float32), scales=np.zeros((N, 3), np.float32),
                 rotations=np.zeros((N, 4), np.float32), cov3D_precomp=np.zeros((N, 6), np.float32))
        acc = np.zeros((max(N, 1), 10), np.float64) if return_acc2d else None
        _L().gso_backward(self._h, _p(g_color), _p(g_depth), _p(g_alpha),
                          _p(g["means3D"]), _p(g["means2D"]), _p(g["shs"]), _p(g["colors_precomp"]), _p(g["opacities"]),
                          _p(g["scales"]), _p(g["rotations"]), _p(g["cov3D_precomp"]), _p(acc), int(bool(f32_replay)))
        if not self.use_sh:
            g.pop("shs")
        else:
            g.pop("colors_precomp")
        if not self.use_cov:
            g.pop("cov3D_precomp")
        else:
            g.pop("scales"), g.pop("rotations")
        if return_acc2d:
            g["acc2d"] = acc[:N]
        return g


def forward(means3D, opacities, image_height, image_width, tanfovx, tanfovy, bg, viewmatrix, projmatrix,
            campos, sh_degree=0, scale_modifier=1.0, shs=None, colors_precomp=None, scales=None,
            rotations=None, cov3D_precomp=None):
    """One view.  Returns (color[3,H,W], radii[N] i32, depth[1,H,W], alpha[1,H,W], state)."""
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    means3D = _f32(means3D).reshape(-1, 3)
    N = means3D.shape[0]
    H, W = int(image_height), int(image_width)
    shs = _f32(shs)
    M = 0 if shs is None else shs.reshape(N, -1, 3).shape[1]
    colors_precomp = _f32(colors_precomp)
    opacities = _f32(opacities).reshape(-1)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    view = _f32(viewmatrix).reshape(16)
    proj = _f32(projmatrix).reshape(16)
    campos = _f32(campos).reshape(3)
    bg = _f32(bg).reshape(3)
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    alpha = np.zeros((1, H, W), np.float32)
    radii = np.zeros(max(N, 1), np.int32)
    h = _L().gso_forward(N, M, int(sh_degree), H, W, _p(means3D), _p(shs), _p(colors_precomp), _p(opacities),
                         _p(scales), _p(rotations), _p(cov3D_precomp), C.c_float(scale_modifier), _p(view), _p(proj),
                         _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), _p(bg),
                         _p(color), _p(depth), _p(alpha), _p(radii))
    st = GsOracleState(h, N, M, H, W, shs is not None, cov3D_precomp is not None)
    return color, radii[:N], depth, alpha, st


def mark_visible(means3D, viewmatrix):
    means3D = _f32(means3D).reshape(-1, 3)
    vis = np.zeros(max(means3D.shape[0], 1), np.uint8)
    _L().gso_mark_visible(means3D.shape[0], _p(means3D), _p(_f32(viewmatrix).reshape(16)), _p(vis))
    return vis[:means3D.shape[0]].astype(bool)
