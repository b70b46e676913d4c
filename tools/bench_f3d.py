"""BASELINE config 4: 8K texture bake (face3d render_colors), GPU kernel vs the reference's own C++ on the host.
    python tools/bench_f3d.py [--grid 245] [--res 8192] [--no-cpu]
Prints one JSON line: device time of f3d_render_colors (inputs resident), end-to-end time through the
NumPy-facing drop-in (H2D + kernels + D2H), algorithmic HBM bytes and fraction of the measured copy peak,
and the CPU reference time (oracle/_ref when present, else the port).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from topo4d_b200 import synth  # noqa: E402
from topo4d_b200.face3d_compat import mesh_core_cython as mcc  # noqa: E402
from topo4d_b200.face3d_compat import render as f3d_render  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=245)
    ap.add_argument("--res", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    res = a.res
    v, t, c = synth.uv_grid_mesh(grid=a.grid, res=res, seed=0)
    dev = torch.device("cuda:0")
    d_v = torch.tensor(v, dtype=torch.float32, device=dev)
    d_t = torch.tensor(t, dtype=torch.int32, device=dev)
    d_c = torch.tensor(c, dtype=torch.float32, device=dev)
    img = torch.zeros((res, res, 3), device=dev)
    dep = torch.full((res, res), -999999.0, device=dev)
    ws = mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.iters):
        img.zero_(); dep.fill_(-999999.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3, ws)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    # end to end through the reference-facing NumPy API (fresh arrays each call, like helpers.py:956)
    t0 = time.perf_counter()
    out = f3d_render.render_colors(v, t, c, res, res, 3)
    e2e_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    out_u8 = f3d_render.render_colors_u8(v, t, c, res, res, 3)
    e2e_u8_s = time.perf_counter() - t0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = res * res * 3 * 4 + v.shape[0] * 24 + t.shape[0] * 12            # image write + mesh read
    line = {"metric": "face3d render_colors 8K bake", "res": res, "triangles": int(t.shape[0]), "vertices": int(v.shape[0]),
            "gpu_ms": ms, "gpu_mpix_s": res * res / 1e6 / (ms / 1e3), "e2e_numpy_s": e2e_s, "e2e_u8_bake_s": e2e_u8_s,
            "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "peak_gbs": peak,
            "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "covered_frac": float((out.sum(-1) != 0).mean())}
    if not a.no_cpu:
        from oracle import f3d_oracle
        fn, kind = (f3d_oracle.render_colors_ref, "reference") if f3d_oracle.have_ref() else (f3d_oracle.render_colors_port, "port")
        t0 = time.perf_counter()
        ref, _ = fn(v, t, c, res, res, 3)
        line["cpu_s"] = time.perf_counter() - t0
        line["cpu_kind"] = kind
        line["cpu_mpix_s"] = res * res / 1e6 / line["cpu_s"]
        line["bit_exact_vs_cpu"] = bool(np.array_equal(ref, out))
        line["u8_matches_cpu"] = bool(np.array_equal((ref * 255).astype(np.uint8), out_u8))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
