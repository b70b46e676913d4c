"""The reference's OWN regimes and BASELINE config 4 as bench.py side workloads (do not over-fit 60 k / 1080p).

  geometry  8 280 mesh-bound Gaussians (one per vertex, train.py:138-146) at 512 x 375 (--down_ratio 8, helpers.py:807),
            ONE view per optimiser step (train.py:105-112,661-673), colors_precomp, opacity 1: eager and CUDA-graph replay
  texture   4 M UV-densified Gaussians at 4096 x 3000 (helpers.py:608-609, train.py:596,715-743), one view per step
  bake      face3d render_colors at 8192^2, 120 050 and 9.59 M triangles (config 4; helpers.py:953-960)
  train     a short run of the config-5 stand-in (tools/train_synthetic.py: the reference-shaped per-frame loop on a synthetic
            24-view sequence read from JPEG files through the frame prefetcher, fused loss / Adam, per-frame bake)

Each returns a plain dict that bench.py attaches under "workloads" on its one JSON line.  Device-timed with CUDA events
after warm-up; inputs resident unless a key says otherwise.  Nothing here imports oracle/ except the bake's CPU column
(the reference's own C++ built into oracle/_ref, timed beside the GPU number like bench.py's cpu_baseline)."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

from topo4d_b200 import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:  # noqa: BLE001
        return 6650.0, "6650 GB/s (of fallback)"


def _loss_grads(color, depth, alpha, gen):
    H, W = color.shape[-2:]
    target = torch.rand(color.shape, device=color.device, generator=gen)
    return (torch.sign(color - target) / (3 * H * W), torch.full_like(depth, 0.1 / (H * W)), torch.full_like(alpha, 0.1 / (H * W)))


def _render_setup(scene, cams, H, W, dev, per_view=True):
    t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
    cam_all = torch.tensor(engine.pack_cameras_numpy(cams, (0.0, 0.0, 0.0)), device=dev)
    views = [cam_all[i:i + 1].contiguous() for i in range(len(cams))] if per_view else [cam_all]
    gen = torch.Generator(device=dev).manual_seed(4321)
    kw = dict(colors_precomp=t.get("colors_precomp"), shs=t.get("shs"), scales=t["scales"], rotations=t["rotations"],
              sh_degree=0 if "shs" not in t else int(round(t["shs"].shape[1] ** 0.5)) - 1)
    caps, gimgs, stats = [], [], []
    for cam in views:
        color, radii, depth, alpha, st = engine.forward(t["means3D"], t["opacities"], cam, H, W, **kw)
        s = st.status()
        caps.append(int(s.num_instances * 1.1) + 4096)
        stats.append((int(s.num_instances), int(s.max_tile_instances), int(s.num_active_tiles)))
        gimgs.append(_loss_grads(color, depth, alpha, gen))
        del color, depth, alpha, st
    return t, views, kw, caps, gimgs, stats


def geometry(steps: int = 240, device: str = "cuda:0") -> dict:
    dev = torch.device(device)
    H, W, N, V = 375, 512, 8280, 24
    scene = synth.head_scene(N, seed=0, sh_degree=None, opacity="topo4d")
    cams = synth.ring_cameras(V, w=W, h=H, radius=0.6, focal_over_h=1.6)
    t, views, kw, caps, gimgs, stats = _render_setup(scene, cams, H, W, dev)
    flat = [None]

    def it(i):
        *_, st = engine.forward(t["means3D"], t["opacities"], views[i], H, W, check="none", cap_instances=caps[i], **kw)
        flat[0] = engine.backward(st, *gimgs[i], flat=flat[0]).flat

    for i in range(V):
        it(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        it(k % V)
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / steps
    # CUDA-graph replay: one graph holds one pass over the 24 cameras (24 optimiser-step renders)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for i in range(V):
            it(i)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for i in range(V):
            it(i)
    reps = max(steps // V, 3)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / (reps * V)
    return {"what": f"reference geometry regime: {N} Gaussians, {W}x{H}, colors_precomp, opacity 1, ONE view per step (fwd+bwd), 24 cameras in turn",
            "ms_per_step_eager": eager_ms, "ms_per_step_graph": graph_ms, "mpix_s_graph": H * W / 1e6 / (graph_ms / 1e3),
            "steps": steps, "num_rendered_per_view": float(np.mean([s[0] for s in stats])), "non_empty_tiles_per_view": float(np.mean([s[2] for s in stats])),
            "kernels_per_step": engine.KERNELS_PER_FORWARD + engine.KERNELS_PER_BACKWARD}


def texture(n: int = 4_000_000, steps: int = 5, device: str = "cuda:0") -> dict:
    dev = torch.device(device)
    H, W = 3000, 4096
    scene = synth.dense_head_scene(n, seed=0)
    cams = synth.ring_cameras(24, w=W, h=H)[:1]
    t, views, kw, caps, gimgs, stats = _render_setup(scene, cams, H, W, dev)
    flat = [None]
    ev = {}

    def it(stage=None):
        *_, st = engine.forward(t["means3D"], t["opacities"], views[0], H, W, check="none", cap_instances=caps[0], stage_events=stage, **kw)
        flat[0] = engine.backward(st, *gimgs[0], flat=flat[0], stage_events=stage).flat
        return st

    for _ in range(2):
        st = it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        it()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = it(ev)
    torch.cuda.synchronize()
    assert not st.status().overflow
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev.items()}
    I, mx, act = stats[0]
    peak, src = _peak()
    P = H * W
    alg = 56 * P + 152 * I + (268 + 3 * 12) * n          # per-view algorithmic bytes, S = 12 (colors_precomp)
    return {"what": f"reference texture regime: {n} Gaussians, {W}x{H}, colors_precomp, one view per step (fwd+bwd)",
            "ms_per_step": ms, "mpix_s": P / 1e6 / (ms / 1e3), "steps": steps, "num_rendered": I, "max_tile_instances": mx,
            "non_empty_tiles": act, "stage_ms": stage_ms, "algorithmic_bytes_per_step": alg,
            "frac_of_hbm_peak_whole_step": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": src}


def bake(grids=(245, 2190), res: int = 8192, iters: int = 5, cpu: bool = True, device: str = "cuda:0") -> dict:
    from topo4d_b200.face3d_compat import mesh_core_cython as mcc
    from topo4d_b200.face3d_compat import render as f3d_render
    dev = torch.device(device)
    peak, src = _peak()
    out = {}
    for grid in grids:
        v, tri, c = synth.uv_grid_mesh(grid=grid, res=res, seed=0)
        d_v = torch.tensor(v, dtype=torch.float32, device=dev)
        d_t = torch.tensor(tri, dtype=torch.int32, device=dev)
        d_c = torch.tensor(c, dtype=torch.float32, device=dev)
        img = torch.zeros((res, res, 3), device=dev)
        dep = torch.full((res, res), -999999.0, device=dev)
        ws = mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3)
        torch.cuda.synchronize()
        times = []
        for _ in range(iters):
            img.zero_(); dep.fill_(-999999.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3, ws)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        del img, dep, ws
        # the bake as face3d/mesh/render.py performs it (fresh image, private depth): fused fill, no depth plane traffic
        bt = {}
        for u8 in (False, True):
            ws = None
            tt = []
            for _ in range(iters + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = mcc.bake_colors_device(d_v, d_t, d_c, res, res, 3, u8=u8)
                e1.record()
                torch.cuda.synchronize()
                tt.append(e0.elapsed_time(e1))
                del o
            bt[u8] = float(np.median(tt[1:]))
        # through the NumPy-facing drop-in: the first full-size call also allocates the page-locked result block (kept by
        # PyTorch's caching host allocator); a per-frame bake (train.py:755: once per frame) runs at the steady-state figure
        t0 = time.perf_counter()
        res_np = f3d_render.render_colors(v, tri, c, res, res, 3)
        first_s = time.perf_counter() - t0
        del res_np
        t0 = time.perf_counter()
        res_np = f3d_render.render_colors(v, tri, c, res, res, 3)
        e2e_s = time.perf_counter() - t0
        res_u8 = f3d_render.render_colors_u8(v, tri, c, res, res, 3)
        del res_u8
        t0 = time.perf_counter()
        res_u8 = f3d_render.render_colors_u8(v, tri, c, res, res, 3)
        e2e_u8_s = time.perf_counter() - t0
        alg = res * res * 3 * 4 + v.shape[0] * 24 + tri.shape[0] * 12            # image write + mesh read
        alg_u8 = res * res * 3 + v.shape[0] * 24 + tri.shape[0] * 12
        r = {"triangles": int(tri.shape[0]), "gpu_ms_inplace_api": ms, "gpu_ms": bt[False], "gpu_ms_u8": bt[True],
             "mpix_s": res * res / 1e6 / (bt[False] / 1e3), "algorithmic_bytes": alg,
             "frac_of_hbm_peak": alg / (bt[False] * 1e-3) / 1e9 / peak, "frac_of_hbm_peak_u8": alg_u8 / (bt[True] * 1e-3) / 1e9 / peak,
             "e2e_numpy_api_s": e2e_s, "e2e_numpy_api_first_call_s": first_s, "e2e_u8_api_s": e2e_u8_s}
        if cpu:
            from oracle import f3d_oracle
            fn, kind = (f3d_oracle.render_colors_ref, "reference") if f3d_oracle.have_ref() else (f3d_oracle.render_colors_port, "port")
            t0 = time.perf_counter()
            ref, _ = fn(v, tri, c, res, res, 3)
            r.update(cpu_s=time.perf_counter() - t0, cpu_kind=kind, cpu_cores=1, bit_exact_vs_cpu=bool(np.array_equal(ref, res_np)),
                     u8_matches_cpu=bool(np.array_equal((ref * 255).astype(np.uint8), res_u8)))
            r["speedup_numpy_api_vs_cpu"] = r["cpu_s"] / e2e_s
            del ref
        del res_np, res_u8
        out[f"{tri.shape[0]}_tris"] = r
    return {"what": f"BASELINE config 4: face3d render_colors at {res}x{res}, c = 3 (gpu_ms: the bake as render.py performs it, inputs "
                    "resident, kernels only; gpu_ms_inplace_api: the in-place image + depth contract of render_colors_core; e2e_*: through "
                    "the NumPy-facing drop-in with host arrays, steady state)", "peak_source": src, **out}


def train(frames: int = 4, iters: int = 100) -> dict:
    """BASELINE config 5 in miniature (the 100-frame, 8-GPU run is profiles/r02k_config5_g8.json): one process, a few frames."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "train.json")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_synthetic.py"), "--frames", str(frames), "--iters", str(iters),
                            "--files", os.path.join(d, "seq"), "--bake", "1024", "--json", out], capture_output=True, text=True, cwd=ROOT,
                           timeout=600)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        rep = json.load(open(out))
    fr = rep["frames"]
    return {"what": f"config-5 stand-in, {frames} frames x {iters} one-view optimiser steps, 8280 Gaussians, 512x375, 24 cameras, ground truth from "
                    "JPEG files via FramePrefetcher, image loss + FusedAdam + per-frame 1024^2 bake (tools/train_synthetic.py)",
            "steps_per_s": rep["steps_per_s"], "ms_per_step_last_frame": fr[-1]["ms_per_step"], "psnr_db_last_frame": fr[-1]["psnr_db"],
            "mean_vertex_err_m_last_frame": fr[-1]["mean_vertex_err_m"], "bake_ms_last_frame": fr[-1].get("bake_ms"),
            "loss_first": fr[0]["loss_first"], "loss_last": fr[-1]["loss_last"]}
