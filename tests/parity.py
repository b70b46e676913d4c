"""Shared parity harness: run the CUDA path (through the C ABI) and the CPU oracle on the same
seeded inputs and report stage-by-stage differences.  Used by the `-m gpu` tests,
`__graft_entry__.smoke()` and tools/gpu_report.py.  (Oracle use is confined to checkers.)

Tolerances (BASELINE.json north_star): RGB / depth / alpha within 1e-4 abs; gradients within
1e-3 relative; radii, tile ranges and per-tile sorted Gaussian lists bit-exact.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import gs_oracle
from topo4d_b200 import engine

ABS_TOL = 1e-4      # colour / depth / alpha
REL_TOL = 1e-3      # gradients


def grad_rel_err(ours: np.ndarray, ref: np.ndarray) -> float:
    """max |ours-ref| / (|ref| + floor): elementwise relative error with a floor of 1e-3 of the tensor's
    largest magnitude, so exact zeros / cancellation noise do not blow the ratio up."""
    scale = float(np.abs(ref).max())
    if scale == 0.0:
        return float(np.abs(ours).max())
    return float((np.abs(ours - ref) / (np.abs(ref) + 1e-3 * scale)).max())


def run_oracle(scene, cams, H, W, sh_degree, bg, g_color=None, g_depth=None, g_alpha=None, scale_modifier=1.0):
    """Per-view oracle forward (+ backward summed over views when g_* given)."""
    outs, grads = [], None
    for v, cam in enumerate(cams):
        color, radii, depth, alpha, st = gs_oracle.forward(
            scene["means3D"], scene["opacities"], shs=scene.get("shs"), colors_precomp=scene.get("colors_precomp"),
            scales=scene.get("scales"), rotations=scene.get("rotations"), cov3D_precomp=scene.get("cov3D_precomp"),
            image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.asarray(bg, np.float32),
            viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos, sh_degree=sh_degree,
            scale_modifier=scale_modifier)
        o = dict(color=color, radii=radii, depth=depth, alpha=alpha, state=st)
        if g_color is not None:
            g = st.backward(g_color[v], None if g_depth is None else g_depth[v], None if g_alpha is None else g_alpha[v])
            if grads is None:
                grads = {k: a.astype(np.float64) for k, a in g.items()}
            else:
                for k, a in g.items():
                    grads[k] += a
        outs.append(o)
    return outs, grads


def run_cuda(scene, cams, H, W, sh_degree, bg, g_color=None, g_depth=None, g_alpha=None, scale_modifier=1.0,
             device="cuda:0"):
    dev = torch.device(device)
    t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
    cam_t = torch.tensor(engine.pack_cameras_numpy(cams, bg), device=dev)
    color, radii, depth, alpha, st = engine.forward(
        t["means3D"], t["opacities"], cam_t, H, W, shs=t.get("shs"), colors_precomp=t.get("colors_precomp"),
        scales=t.get("scales"), rotations=t.get("rotations"), cov3D_precomp=t.get("cov3D_precomp"),
        sh_degree=sh_degree, scale_modifier=scale_modifier)
    out = dict(color=color, radii=radii, depth=depth, alpha=alpha, state=st)
    grads = None
    if g_color is not None:
        gb = engine.backward(st, torch.tensor(g_color, device=dev),
                             None if g_depth is None else torch.tensor(g_depth, device=dev),
                             None if g_alpha is None else torch.tensor(g_alpha, device=dev))
        grads = {k: getattr(gb, k) for k in ("means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                             "rotations", "cov3D_precomp") if getattr(gb, k) is not None}
    return out, grads


def compare(scene, cams, H, W, sh_degree=0, bg=(0.0, 0.0, 0.0), seed=1, with_backward=True, scale_modifier=1.0,
            device="cuda:0") -> dict:
    """Returns a flat dict of metrics (see keys below); asserts nothing."""
    V = len(cams)
    rng = np.random.default_rng(seed)
    gC = gD = gA = None
    if with_backward:
        gC = rng.normal(size=(V, 3, H, W)).astype(np.float32)
        gD = rng.normal(size=(V, 1, H, W)).astype(np.float32)
        gA = rng.normal(size=(V, 1, H, W)).astype(np.float32)
    ref, ref_g = run_oracle(scene, cams, H, W, sh_degree, bg, gC, gD, gA, scale_modifier)
    out, g = run_cuda(scene, cams, H, W, sh_degree, bg, gC, gD, gA, scale_modifier, device)
    m = {}
    view = out["state"].view()
    m["num_rendered_ref"] = int(sum(r["state"].num_rendered for r in ref))
    m["num_rendered"] = int(view["num_instances"])
    col = out["color"].cpu().numpy(); dep = out["depth"].cpu().numpy(); alp = out["alpha"].cpu().numpy()
    rad = out["radii"].cpu().numpy()
    m["radii_mismatch"] = int(sum((rad[v] != ref[v]["radii"]).sum() for v in range(V)))
    # tile ranges: our exclusive scan vs the oracle's per-view (start,end)
    T = view["tiles_x"] * view["tiles_y"]
    ts = view["tile_start"].cpu().numpy().astype(np.int64)
    ids = view["sorted_ids"].cpu().numpy()
    range_bad = ids_bad = 0
    off = 0
    for v in range(V):
        b = ref[v]["state"].binning()
        cnt_ref = (b["ranges"][:, 1].astype(np.int64) - b["ranges"][:, 0].astype(np.int64))
        cnt = ts[v * T + 1:(v + 1) * T + 1] - ts[v * T:(v + 1) * T]
        range_bad += int((cnt != cnt_ref).sum())
        n = len(b["ids"])
        seg = ids[off:off + n]
        ids_bad += int((seg != b["ids"].astype(np.int32)).sum()) if len(seg) == n else max(n, 1)
        off += n
    m["tile_range_mismatch"] = range_bad
    m["sorted_id_mismatch"] = ids_bad
    nc = view["n_contrib"].cpu().numpy(); fT = view["final_T"].cpu().numpy()
    nc_bad = 0; fT_err = 0.0
    err_c = err_d = err_a = 0.0
    bad_px = 0
    for v in range(V):
        im = ref[v]["state"].image_state()
        diff_nc = nc[v] != im["n_contrib"].astype(np.int32)
        nc_bad += int(diff_nc.sum())
        fT_err = max(fT_err, float(np.abs(fT[v] - im["final_T"])[~diff_nc].max(initial=0.0)))
        dc = np.abs(col[v] - ref[v]["color"]).max(0); dd = np.abs(dep[v] - ref[v]["depth"])[0]; da = np.abs(alp[v] - ref[v]["alpha"])[0]
        # pixels whose threshold decisions (alpha<1/255, T<1e-4) flipped are reported separately
        ok = ~diff_nc
        err_c = max(err_c, float(dc[ok].max(initial=0.0))); err_d = max(err_d, float(dd[ok].max(initial=0.0)))
        err_a = max(err_a, float(da[ok].max(initial=0.0)))
        bad_px += int(((dc > ABS_TOL) | (dd > ABS_TOL) | (da > ABS_TOL)).sum())
    m["n_contrib_mismatch"] = nc_bad
    m["pixels"] = V * H * W
    m["final_T_maxerr"] = fT_err
    m["color_maxerr"] = err_c; m["depth_maxerr"] = err_d; m["alpha_maxerr"] = err_a
    m["pixels_over_tol"] = bad_px
    if with_backward:
        for k, a in ref_g.items():
            if k == "acc2d":
                continue
            m["grad_relerr_" + k] = grad_rel_err(g[k].cpu().numpy().astype(np.float64).reshape(a.shape), a)
    return m


def assert_parity(m: dict, allow_flips: int = 0):
    assert m["num_rendered"] == m["num_rendered_ref"], m
    assert m["radii_mismatch"] == 0, m
    assert m["tile_range_mismatch"] == 0, m
    assert m["sorted_id_mismatch"] == 0, m
    assert m["n_contrib_mismatch"] <= allow_flips, m
    assert m["pixels_over_tol"] <= allow_flips, m
    assert m["color_maxerr"] <= ABS_TOL and m["depth_maxerr"] <= ABS_TOL and m["alpha_maxerr"] <= ABS_TOL, m
    for k, v in m.items():
        if k.startswith("grad_relerr_"):
            assert v <= REL_TOL, (k, v, m)
