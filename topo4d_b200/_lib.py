"""ctypes binding of libtopo4d_b200.so (the C ABI in include/topo4d_b200.h).

The product path has no CPU fallback: if the CUDA library is missing and cannot be built,
importing this module's :func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

GS_CAM_FLOATS = 48
GS_HINT_UNKNOWN, GS_HINT_SHORT_LISTS, GS_HINT_LONG_LISTS = 0, 1, 2
GS_FWD_PREPROCESS, GS_FWD_SCATTER, GS_FWD_SORT, GS_FWD_BLEND, GS_FWD_ALL = 1, 2, 4, 8, 15
GS_BWD_BLEND, GS_BWD_PREPROCESS, GS_BWD_ALL = 1, 2, 3
GS_OK, GS_E_BAD_ARGS, GS_E_WORKSPACE_SMALL, GS_E_CUDA, GS_E_OVERFLOW, GS_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

_vp = C.c_void_p


class GsProblem(C.Structure):
    _fields_ = [("N", C.c_int32), ("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("scale_modifier", C.c_float),
                ("debug", C.c_int32), ("blend_px", C.c_int32), ("hints", C.c_int32), ("cap_instances", C.c_int64),
                ("means3D", _vp), ("shs", _vp), ("colors_precomp", _vp), ("opacities", _vp), ("scales", _vp),
                ("rotations", _vp), ("cov3D_precomp", _vp), ("cameras", _vp),
                ("workspace", _vp), ("workspace_bytes", C.c_size_t)]


class GsForwardOut(C.Structure):
    _fields_ = [("color", _vp), ("depth", _vp), ("alpha", _vp), ("radii", _vp)]


class GsBackwardIO(C.Structure):
    _fields_ = [("dL_dcolor", _vp), ("dL_ddepth", _vp), ("dL_dalpha", _vp), ("radii", _vp),
                ("dL_dmeans3D", _vp), ("dL_dmeans2D", _vp), ("dL_dshs", _vp), ("dL_dcolors", _vp),
                ("dL_dopacities", _vp), ("dL_dscales", _vp), ("dL_drotations", _vp), ("dL_dcov3D", _vp)]


class GsStatus(C.Structure):
    _fields_ = [("num_instances", C.c_int64), ("cap_instances", C.c_int64), ("overflow", C.c_int32),
                ("max_tile_instances", C.c_int32), ("num_active_tiles", C.c_int32), ("reserved0", C.c_int32)]


class GsWorkspaceView(C.Structure):
    _fields_ = [("tile_start", _vp), ("sorted_ids", _vp), ("sorted_records", _vp), ("geom_records", _vp),
                ("final_T", _vp), ("n_contrib", _vp), ("grad2d", _vp), ("tiles_x", C.c_int32), ("tiles_y", C.c_int32)]


class T4dImageLoss(C.Structure):
    _fields_ = [("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("w_l1", C.c_float), ("w_ssim", C.c_float),
                ("render", _vp), ("target", _vp), ("cam_m", _vp), ("cam_c", _vp), ("loss", _vp), ("dL_drender", _vp),
                ("dL_dcam_m", _vp), ("dL_dcam_c", _vp), ("workspace", _vp), ("workspace_bytes", C.c_size_t)]


T4D_ADAM_MAX_SEGMENTS = 24


class T4dAdamSegment(C.Structure):
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("pin_mask", _vp), ("pin_values", _vp),
                ("count", C.c_int64), ("row_width", C.c_int32), ("step", C.c_int32), ("lr", C.c_float), ("step_device", _vp),
                ("lr_device", _vp)]


# every symbol include/topo4d_b200.h declares: (name, restype, argtypes)
SYMBOLS = {
    "gs_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "gs_forward": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsForwardOut), _vp]),
    "gs_backward": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsBackwardIO), _vp]),
    "gs_forward_stages": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsForwardOut), C.c_uint32, _vp]),
    "gs_backward_stages": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsBackwardIO), C.c_uint32, _vp]),
    "gs_read_status": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsStatus), _vp]),
    "gs_count_instances": (C.c_int, [C.POINTER(GsProblem), C.POINTER(C.c_int64), _vp]),
    "gs_mark_visible": (C.c_int, [C.c_int32, _vp, _vp, _vp, _vp]),
    "gs_workspace_view": (C.c_int, [C.POINTER(GsProblem), C.POINTER(GsWorkspaceView)]),
    "gs_last_error": (C.c_char_p, [C.c_int]),
    "gs_last_cuda_error": (C.c_char_p, []),
    "f3d_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "f3d_render_colors": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    _vp, C.c_size_t, _vp]),
    "f3d_set_band_bytes": (None, [C.c_int64]),
    "f3d_bake_colors": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  _vp, C.c_size_t, _vp]),
    "f3d_render_colors_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "f3d_image_to_u8": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "t4d_image_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "t4d_image_loss": (C.c_int, [C.POINTER(T4dImageLoss), _vp]),
    "t4d_dense_attribute": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_int32, _vp, _vp]),
    "t4d_activate": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "t4d_activate_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "t4d_adam_step": (C.c_int, [C.POINTER(T4dAdamSegment), C.c_int32, C.c_float, C.c_float, C.c_float, _vp]),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB_PATH


def lib():
    """Load the CUDA library, (re)building it in-tree when it is missing or does not match its sources' content hash.
    Raises if that is impossible."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("TOPO4D_B200_LIB") or _build.LIB_PATH      # override: experiment variants (build.py --tag)
        if path == _build.LIB_PATH:
            # build_library() returns at once when the library matches the content hash of its sources; a stale binary
            # (edited .cu / header, ctypes structs out of step) is rebuilt -- by one process only under torchrun
            if int(os.environ.get("LOCAL_RANK", "0")) == 0 or not os.path.exists(path):
                path = _build.build_library(force=os.environ.get("TOPO4D_B200_REBUILD") == "1")
            elif _build._stale(path):
                import warnings
                warnings.warn("topo4d_b200: libtopo4d_b200.so is older than its sources and this is not local rank 0; "
                              "loading it as is (run `python -m topo4d_b200.build`)")
        handle = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)        # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


class GsError(RuntimeError):
    pass


def check(code: int, what: str = "") -> None:
    if code == 0:
        return
    L = lib()
    msg = L.gs_last_error(code).decode()
    if code == GS_E_CUDA:
        msg += " -- " + L.gs_last_cuda_error().decode()
    raise GsError(f"{what}: {msg} (code {code})")
