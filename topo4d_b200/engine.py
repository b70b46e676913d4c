"""Functional host layer over the C ABI: tensors in, tensors out, no autograd.

PyTorch is plumbing here (device memory, streams); every kernel on this path lives in
libtopo4d_b200.so.  One call renders V camera views of the same Gaussians (the reference renders
one view per call -- train.py:307 -- which is V = 1).
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import GS_CAM_FLOATS, GsBackwardIO, GsForwardOut, GsProblem, GsStatus, GsWorkspaceView

FWD_STAGES = (("preprocess", _lib.GS_FWD_PREPROCESS), ("scatter", _lib.GS_FWD_SCATTER), ("sort_gather", _lib.GS_FWD_SORT),
              ("blend_fwd", _lib.GS_FWD_BLEND))
BWD_STAGES = (("blend_bwd", _lib.GS_BWD_BLEND), ("preprocess_bwd", _lib.GS_BWD_PREPROCESS))
# kernels launched by one forward / one backward call (memsets are not kernels): preprocess, fused tile scan (two
# launches only when V*tiles > 4096 * SM count), scatter, sort_gather, blend_fwd | blend_bwd, preprocess_bwd
KERNELS_PER_FORWARD, KERNELS_PER_BACKWARD = 5, 2

# remembered instance capacity per problem shape (grown on overflow)
_CAP_MEMO: dict[tuple, int] = {}
_COUNT_MEMO: dict[tuple, int] = {}  # last measured (tile, Gaussian) instance count per problem shape
_TOUCHED_KEYS: set = set()         # problem shapes rendered since the set was last cleared (graph.capture sizes headroom)
# remembered number of non-empty tiles per problem shape: picks the blend kernels' pixels-per-thread variant
_ACTIVE_MEMO: dict[tuple, int] = {}
_MAXTILE_MEMO: dict[tuple, int] = {}  # longest per-tile list per problem shape: tells the library whether the long-list sort kernel is needed
_WS_BYTES: dict[tuple, int] = {}
ctypes_sizeof_status_dev = 48      # sizeof(GsStatusDev), csrc/gs_common.cuh (layout mirrored in RasterState.status)
_PINNED_RING = None                # one pinned allocation, 256 status slots handed out round-robin
_PINNED_NEXT = 0


_LOCK = threading.Lock()            # guards the pinned ring cursor; the memo dictionaries are keyed per device and only ever
                                    # updated with single dict stores (atomic under the GIL)


def _pinned_slot() -> torch.Tensor:
    global _PINNED_RING, _PINNED_NEXT
    with _LOCK:
        if _PINNED_RING is None:
            _PINNED_RING = torch.empty(256 * 64, dtype=torch.uint8).pin_memory()
        i = _PINNED_NEXT
        _PINNED_NEXT = (i + 1) % 256
    return _PINNED_RING[i * 64:i * 64 + ctypes_sizeof_status_dev]


def pick_blend_px(num_active_tiles: int | None) -> int:
    """4 pixels/thread when there are enough non-empty tiles to fill the GPU (fewest instructions), 2 or 1 when
    there are few (more warps per tile: the per-tile latency bounds such launches).  Never changes results."""
    if num_active_tiles is None or num_active_tiles >= 20000:
        return 4
    return 2 if num_active_tiles >= 1200 else 1


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor | None, dev) -> torch.Tensor | None:
    if t is None:
        return None
    if t.dtype is torch.float32 and t.device == dev:
        t = t if t.is_contiguous() else t.contiguous()
    else:
        t = t.to(device=dev, dtype=torch.float32).contiguous()
    # the kernels use 16-byte vector loads on per-Gaussian rows (rotations, SH): a contiguous slice of a flat parameter buffer
    # at an offset that is not a multiple of four floats would fault with a sticky 'misaligned address'
    return t.clone() if (t.data_ptr() & 15) and t.numel() else t


def pack_cameras_numpy(cams, bg=(0.0, 0.0, 0.0)) -> np.ndarray:
    """List of synth.Camera-like objects (viewmatrix, projmatrix, campos, tanfovx, tanfovy) -> [V,48] float32."""
    out = np.zeros((len(cams), GS_CAM_FLOATS), np.float32)
    for i, c in enumerate(cams):
        out[i, 0:16] = np.asarray(c.viewmatrix, np.float32).reshape(16)
        out[i, 16:32] = np.asarray(c.projmatrix, np.float32).reshape(16)
        out[i, 32:35] = np.asarray(c.campos, np.float32).reshape(3)
        out[i, 35:38] = np.asarray(bg, np.float32).reshape(3)
        out[i, 38] = c.tanfovx
        out[i, 39] = c.tanfovy
    return out


@dataclass
class RasterState:
    """Everything backward needs; owns the workspace so concurrent forwards never alias."""
    problem: GsProblem
    workspace: torch.Tensor
    keep: tuple                      # input tensors kept alive (their pointers sit in `problem`)
    radii: torch.Tensor
    N: int
    V: int
    H: int
    W: int
    M: int
    use_sh: bool
    use_cov: bool
    device: torch.device
    key: tuple = ()
    _status: GsStatus | None = field(default=None, repr=False)
    _status_pinned: torch.Tensor | None = field(default=None, repr=False)     # async copy of the status block ...
    _status_event: torch.cuda.Event | None = field(default=None, repr=False)  # ... complete once this event has fired

    def prefetch_status(self) -> None:
        """Enqueue an asynchronous copy of the status block into pinned memory right behind the forward: a later
        status() then only waits for THAT point of the stream, not for whatever was queued afterwards."""
        self._status_pinned = _pinned_slot()
        self._status_pinned.copy_(self.workspace[:ctypes_sizeof_status_dev], non_blocking=True)
        self._status_event = torch.cuda.Event()
        self._status_event.record(torch.cuda.current_stream(self.device))

    def status(self) -> GsStatus:
        """Synchronising read of the device status block (num_rendered, overflow, longest tile list)."""
        if self._status is None and self._status_pinned is not None:
            self._status_event.synchronize()
            raw = self._status_pinned.numpy()
            st = GsStatus()
            st.num_instances = int(raw[0:8].view("<i8")[0])
            st.cap_instances = int(raw[8:16].view("<i8")[0])
            st.overflow = int(raw[16:20].view("<i4")[0])
            st.max_tile_instances = int(raw[20:24].view("<i4")[0])
            st.num_active_tiles = int(raw[24:28].view("<u4")[0]) + int(raw[44:48].view("<u4")[0])   # num_long + num_short
            self._status = st
            if self.key:
                _ACTIVE_MEMO[self.key] = int(st.num_active_tiles)
                _COUNT_MEMO[self.key] = int(st.num_instances)
                _MAXTILE_MEMO[self.key] = int(st.max_tile_instances)
        if self._status is None:
            st = GsStatus()
            with torch.cuda.device(self.device):
                code = _lib.lib().gs_read_status(C.byref(self.problem), C.byref(st),
                                                 C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if code not in (_lib.GS_OK, _lib.GS_E_OVERFLOW):
                _lib.check(code, "gs_read_status")
            self._status = st
            if self.key:
                _ACTIVE_MEMO[self.key] = int(st.num_active_tiles)
                _COUNT_MEMO[self.key] = int(st.num_instances)
                _MAXTILE_MEMO[self.key] = int(st.max_tile_instances)
        return self._status

    def view(self) -> dict:
        """Index structures / per-pixel state as tensors (clones), for the bit-exact parity tests."""
        wv = GsWorkspaceView()
        _lib.check(_lib.lib().gs_workspace_view(C.byref(self.problem), C.byref(wv)), "gs_workspace_view")
        base = self.workspace.data_ptr()
        T = wv.tiles_x * wv.tiles_y * self.V
        I = min(int(self.status().num_instances), int(self.problem.cap_instances))

        def sl(ptr, nbytes, dtype):
            off = ptr - base
            return self.workspace[off:off + nbytes].view(dtype).clone()
        VN = self.V * max(self.N, 1)
        return dict(
            tile_start=sl(wv.tile_start, 4 * (T + 1), torch.int32),
            sorted_ids=sl(wv.sorted_ids, 4 * I, torch.int32),
            sorted_records=sl(wv.sorted_records, 48 * I, torch.float32).view(-1, 12),
            geom_records=sl(wv.geom_records, 48 * VN, torch.float32).view(self.V, -1, 12),
            final_T=sl(wv.final_T, 4 * self.V * self.H * self.W, torch.float32).view(self.V, self.H, self.W),
            n_contrib=sl(wv.n_contrib, 4 * self.V * self.H * self.W, torch.int32).view(self.V, self.H, self.W),
            grad2d=sl(wv.grad2d, 48 * VN, torch.float32).view(self.V, -1, 12),
            tiles_x=wv.tiles_x, tiles_y=wv.tiles_y, num_instances=int(self.status().num_instances))


def _check_pairs(shs, colors_precomp, scales, rotations, cov3D_precomp):
    # same messages (typo included) as the upstream Python wrapper the reference imports (train.py:19)
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')


def forward(means3D, opacities, cameras, image_height, image_width, *, shs=None, colors_precomp=None,
            scales=None, rotations=None, cov3D_precomp=None, sh_degree=0, scale_modifier=1.0, debug=False,
            check="sync", cap_instances=None, stage_events=None, blend_px=None):
    """Render V views.  `cameras`: [V,48] float32 CUDA tensor (see include/topo4d_b200.h GS_CAM_*).

    check: "sync"  -> read the status block after the launch (one small D2H, like upstream's
                      num_rendered read) and transparently re-run with a larger capacity on overflow;
           "none"  -> fully asynchronous; call ``state.status()`` later to validate (synchronises the stream);
           "deferred" -> asynchronous, and the status block is copied to pinned memory right behind the forward
                      so that a later ``state.status()`` waits for this forward only.
    stage_events: optional dict; when given, the pipeline is issued one stage at a time
           (gs_forward_stages) with CUDA events recorded around every stage on the current stream
           and appended to stage_events[stage_name] as (start, end) pairs.
    Returns (color[V,3,H,W], radii[V,N] i32, depth[V,1,H,W], alpha[V,1,H,W], RasterState)."""
    _check_pairs(shs, colors_precomp, scales, rotations, cov3D_precomp)
    if not means3D.is_cuda:
        raise RuntimeError("topo4d_b200: the rasterizer is CUDA-only (means3D is on %s); there is no CPU path" % means3D.device)
    L = _lib.lib()
    dev = means3D.device
    H, W = int(image_height), int(image_width)
    means3D = _f32c(means3D, dev)
    N = int(means3D.shape[0])
    opacities = _f32c(opacities, dev)
    shs = _f32c(shs, dev)
    colors_precomp = _f32c(colors_precomp, dev)
    scales, rotations, cov3D_precomp = _f32c(scales, dev), _f32c(rotations, dev), _f32c(cov3D_precomp, dev)
    cameras = _f32c(cameras, dev).reshape(-1, GS_CAM_FLOATS)
    V = int(cameras.shape[0])
    M = 0 if shs is None else int(shs.reshape(N, -1, 3).shape[1]) if N > 0 else int(shs.shape[1])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    # one allocation for the three images (each stays a contiguous tensor of its own shape)
    P = H * W
    imgs = torch.empty(V * 5 * P, dtype=torch.float32, device=dev)
    color = imgs[:V * 3 * P].view(V, 3, H, W)
    depth = imgs[V * 3 * P:V * 4 * P].view(V, 1, H, W)
    alpha = imgs[V * 4 * P:].view(V, 1, H, W)
    radii = torch.empty((V, max(N, 1)), dtype=torch.int32, device=dev)
    out = GsForwardOut(_ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii))

    key = (N, V, H, W, dev.index)
    _TOUCHED_KEYS.add(key)
    cap = int(cap_instances) if cap_instances is not None else _CAP_MEMO.get(key)
    px = int(blend_px) if blend_px else pick_blend_px(_ACTIVE_MEMO.get(key))
    mt = _MAXTILE_MEMO.get(key)
    # lists a little under the 2048-key limit may grow past it while the splats move: keep the long-list kernel from 1536 up
    hints = _lib.GS_HINT_UNKNOWN if mt is None else (_lib.GS_HINT_SHORT_LISTS if mt <= 1536 else _lib.GS_HINT_LONG_LISTS)

    def make_problem(cap_):
        wk = (N, V, H, W, cap_)
        nbytes = _WS_BYTES.get(wk)
        if nbytes is None:
            nbytes = _WS_BYTES[wk] = L.gs_workspace_bytes(N, V, H, W, cap_)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pr = GsProblem(N, V, H, W, int(sh_degree), M, float(scale_modifier), int(bool(debug)), px, hints, cap_,
                       _ptr(means3D), _ptr(shs), _ptr(colors_precomp), _ptr(opacities), _ptr(scales), _ptr(rotations),
                       _ptr(cov3D_precomp), _ptr(cameras), _ptr(ws), nbytes)
        return pr, ws

    with torch.cuda.device(dev):
        if cap is None and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("topo4d_b200: the first render of a new shape sizes its workspace with a synchronising counting "
                               "pass and cannot be captured in a CUDA graph: run the step eagerly once before capturing")
        if cap is None:
            # first call for this shape: exact count from a preprocess-only pass
            pr, ws = make_problem(0)
            n = C.c_int64(0)
            _lib.check(L.gs_count_instances(C.byref(pr), C.byref(n), stream), "gs_count_instances")
            cap = int(n.value * (1.25 if check == "sync" else 2.0)) + 4096
        while True:
            pr, ws = make_problem(cap)
            if stage_events is None:
                _lib.check(L.gs_forward(C.byref(pr), C.byref(out), stream), "gs_forward")
            else:
                for name, bit in FWD_STAGES:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(L.gs_forward_stages(C.byref(pr), C.byref(out), bit, stream), "gs_forward_stages:" + name)
                    e1.record()
                    stage_events.setdefault(name, []).append((e0, e1))
            state = RasterState(pr, ws, (means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp, cameras),
                                radii, N, V, H, W, M, shs is not None, cov3D_precomp is not None, dev, key)
            if check == "deferred":
                state.prefetch_status()
            if check != "sync":
                break
            st = state.status()
            if not st.overflow:
                break
            cap = int(st.num_instances * 1.25) + 4096
    if cap_instances is None:
        _CAP_MEMO[key] = max(cap, _CAP_MEMO.get(key, 0))
    return color, radii[:, :N], depth, alpha, state


@dataclass
class GradBundle:
    """Gradients of one backward: `flat` is ONE contiguous fp32 buffer (all-reduce-ready), the named
    tensors are views into it."""
    flat: torch.Tensor
    means3D: torch.Tensor
    means2D: torch.Tensor
    opacities: torch.Tensor
    shs: torch.Tensor | None = None
    colors_precomp: torch.Tensor | None = None
    scales: torch.Tensor | None = None
    rotations: torch.Tensor | None = None
    cov3D_precomp: torch.Tensor | None = None


def _al4(n: int) -> int:
    return (n + 3) // 4 * 4


def flat_layout(N: int, M: int, use_sh: bool, use_cov: bool):
    """Segment offsets (in floats, each 16-byte aligned) of the flat gradient buffer: means3D | means2D | colour or SH |
    opacities | scales or cov3D | rotations.  Returns ({name: (offset, size)}, total)."""
    ncol = N * M * 3 if use_sh else N * 3
    sizes = [("means3D", N * 3), ("means2D", N * 3), ("color", ncol), ("opacities", N),
             ("a", N * 6 if use_cov else N * 3), ("b", 0 if use_cov else N * 4)]
    offs, o = {}, 0
    for name, n in sizes:
        offs[name] = (o, n)
        o += _al4(n)
    return offs, max(o, 4)


def flat_views(flat: torch.Tensor, N: int, M: int, use_sh: bool, use_cov: bool) -> dict:
    """Named views of a flat gradient buffer laid out by :func:`flat_layout` (keys as the op's keyword arguments)."""
    offs, _ = flat_layout(N, M, use_sh, use_cov)
    seg = {k: flat[s:s + n] for k, (s, n) in offs.items()}
    out = {"means3D": seg["means3D"].view(N, 3), "means2D": seg["means2D"].view(N, 3), "opacities": seg["opacities"].view(N, 1)}
    out["shs" if use_sh else "colors_precomp"] = seg["color"].view(N, M, 3) if use_sh else seg["color"].view(N, 3)
    if use_cov:
        out["cov3D_precomp"] = seg["a"].view(N, 6)
    else:
        out["scales"], out["rotations"] = seg["a"].view(N, 3), seg["b"].view(N, 4)
    return out


def backward(state: RasterState, dL_dcolor, dL_ddepth=None, dL_dalpha=None, flat: torch.Tensor | None = None,
             stage_events=None) -> GradBundle:
    """Gradients summed over the V views of `state`.  dL_dcolor [V,3,H,W] (or [3,H,W] when V == 1)."""
    L = _lib.lib()
    dev, N, V, H, W, M = state.device, state.N, state.V, state.H, state.W, state.M
    dL_dcolor = _f32c(dL_dcolor, dev)
    dL_ddepth = _f32c(dL_ddepth, dev)
    dL_dalpha = _f32c(dL_dalpha, dev)
    assert dL_dcolor.numel() == V * 3 * H * W, "dL_dcolor must be [V,3,H,W]"
    offs, total = flat_layout(N, M, state.use_sh, state.use_cov)
    if flat is None:
        flat = torch.zeros(total, dtype=torch.float32, device=dev) if N == 0 else \
            torch.empty(total, dtype=torch.float32, device=dev)
    else:
        assert flat.numel() >= total and flat.is_contiguous() and flat.device == dev
    seg = {k: flat[s:s + n] for k, (s, n) in offs.items()}
    io = GsBackwardIO(_ptr(dL_dcolor), _ptr(dL_ddepth), _ptr(dL_dalpha), _ptr(state.radii),
                      _ptr(seg["means3D"]), _ptr(seg["means2D"]),
                      _ptr(seg["color"]) if state.use_sh else None, None if state.use_sh else _ptr(seg["color"]),
                      _ptr(seg["opacities"]),
                      None if state.use_cov else _ptr(seg["a"]), None if state.use_cov else _ptr(seg["b"]),
                      _ptr(seg["a"]) if state.use_cov else None)
    if N > 0:
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if stage_events is None:
                _lib.check(L.gs_backward(C.byref(state.problem), C.byref(io), stream), "gs_backward")
            else:
                for name, bit in BWD_STAGES:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(L.gs_backward_stages(C.byref(state.problem), C.byref(io), bit, stream), "gs_backward_stages:" + name)
                    e1.record()
                    stage_events.setdefault(name, []).append((e0, e1))
    g = GradBundle(flat=flat, means3D=seg["means3D"].view(N, 3), means2D=seg["means2D"].view(N, 3),
                   opacities=seg["opacities"].view(N, 1))
    if state.use_sh:
        g.shs = seg["color"].view(N, M, 3)
    else:
        g.colors_precomp = seg["color"].view(N, 3)
    if state.use_cov:
        g.cov3D_precomp = seg["a"].view(N, 6)
    else:
        g.scales = seg["a"].view(N, 3)
        g.rotations = seg["b"].view(N, 4)
    return g


def mark_visible(positions: torch.Tensor, camera: torch.Tensor) -> torch.Tensor:
    dev = positions.device
    if not positions.is_cuda:
        raise RuntimeError("topo4d_b200: markVisible is CUDA-only")
    pos = _f32c(positions, dev)
    cam = _f32c(camera, dev).reshape(-1)[:GS_CAM_FLOATS]
    vis = torch.empty(pos.shape[0], dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gs_mark_visible(pos.shape[0], _ptr(pos), _ptr(cam), _ptr(vis),
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "gs_mark_visible")
    return vis.bool()
