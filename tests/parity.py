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


def oracle_forward(scene, cams, H, W, sh_degree, bg, scale_modifier=1.0):
    outs = []
    for cam in cams:
        color, radii, depth, alpha, st = gs_oracle.forward(
            scene["means3D"], scene["opacities"], shs=scene.get("shs"), colors_precomp=scene.get("colors_precomp"),
            scales=scene.get("scales"), rotations=scene.get("rotations"), cov3D_precomp=scene.get("cov3D_precomp"),
            image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.asarray(bg, np.float32),
            viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos, sh_degree=sh_degree,
            scale_modifier=scale_modifier)
        outs.append(dict(color=color, radii=radii, depth=depth, alpha=alpha, state=st))
    return outs


def oracle_backward(outs, g_color, g_depth=None, g_alpha=None, f32_replay=False):
    """Oracle gradients summed over the views (float64 accumulation).  f32_replay: the reference-faithful
    all-fp32 blend replay instead of the accurate (double) one -- the fp32 algorithm's own noise floor."""
    grads = None
    for v, o in enumerate(outs):
        g = o["state"].backward(g_color[v], None if g_depth is None else g_depth[v], None if g_alpha is None else g_alpha[v],
                                f32_replay=f32_replay)
        if grads is None:
            grads = {k: a.astype(np.float64) for k, a in g.items()}
        else:
            for k, a in g.items():
                grads[k] += a
    return grads


def run_oracle(scene, cams, H, W, sh_degree, bg, g_color=None, g_depth=None, g_alpha=None, scale_modifier=1.0):
    """Per-view oracle forward (+ backward summed over views when g_* given)."""
    outs = oracle_forward(scene, cams, H, W, sh_degree, bg, scale_modifier)
    return outs, (None if g_color is None else oracle_backward(outs, g_color, g_depth, g_alpha))


def cuda_forward(scene, cams, H, W, sh_degree, bg, scale_modifier=1.0, device="cuda:0", blend_px=None):
    dev = torch.device(device)
    t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
    cam_t = torch.tensor(engine.pack_cameras_numpy(cams, bg), device=dev)
    color, radii, depth, alpha, st = engine.forward(
        t["means3D"], t["opacities"], cam_t, H, W, shs=t.get("shs"), colors_precomp=t.get("colors_precomp"),
        scales=t.get("scales"), rotations=t.get("rotations"), cov3D_precomp=t.get("cov3D_precomp"),
        sh_degree=sh_degree, scale_modifier=scale_modifier, blend_px=blend_px)
    return dict(color=color, radii=radii, depth=depth, alpha=alpha, state=st)


def cuda_backward(out, g_color, g_depth=None, g_alpha=None):
    dev = out["color"].device
    gb = engine.backward(out["state"], torch.tensor(g_color, device=dev),
                         None if g_depth is None else torch.tensor(g_depth, device=dev),
                         None if g_alpha is None else torch.tensor(g_alpha, device=dev))
    return {k: getattr(gb, k) for k in ("means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp") if getattr(gb, k) is not None}


def run_cuda(scene, cams, H, W, sh_degree, bg, g_color=None, g_depth=None, g_alpha=None, scale_modifier=1.0,
             device="cuda:0"):
    out = cuda_forward(scene, cams, H, W, sh_degree, bg, scale_modifier, device)
    return out, (None if g_color is None else cuda_backward(out, g_color, g_depth, g_alpha))


# Oracle results of the large cases, keyed by the caller's `cache_key`: the full-size config-2 tests run the CUDA path
# several times (blend_px = 1 / 2 / 4) against ONE oracle pass (the oracle is the slow side: seconds per 1080p view).
_ORACLE_CACHE: dict = {}


def compare(scene, cams, H, W, sh_degree=0, bg=(0.0, 0.0, 0.0), seed=1, with_backward=True, scale_modifier=1.0,
            device="cuda:0", noise_floor=False, blend_px=None, cache_key=None) -> dict:
    """Returns a flat dict of metrics (see keys below); asserts nothing.

    Pixels whose discrete threshold decisions (alpha < 1/255, T(1-alpha) < 1e-4) flipped between the two
    implementations -- a 1-ulp event in exp() -- are COUNTED (n_contrib_mismatch) and excluded from the value
    comparisons: their loss gradients are zeroed for both sides before the backward passes."""
    V = len(cams)
    rng = np.random.default_rng(seed)
    if cache_key is not None and ("fwd", cache_key) in _ORACLE_CACHE:
        ref = _ORACLE_CACHE[("fwd", cache_key)]
    else:
        ref = oracle_forward(scene, cams, H, W, sh_degree, bg, scale_modifier)
        if cache_key is not None:
            _ORACLE_CACHE[("fwd", cache_key)] = ref
    out = cuda_forward(scene, cams, H, W, sh_degree, bg, scale_modifier, device, blend_px)
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
    flips = np.zeros((V, H, W), bool)
    for v in range(V):
        im = ref[v]["state"].image_state()
        diff_nc = nc[v] != im["n_contrib"].astype(np.int32)
        flips[v] = diff_nc
        nc_bad += int(diff_nc.sum())
        fT_err = max(fT_err, float(np.abs(fT[v] - im["final_T"])[~diff_nc].max(initial=0.0)))
        dc = np.abs(col[v] - ref[v]["color"]).max(0); dd = np.abs(dep[v] - ref[v]["depth"])[0]; da = np.abs(alp[v] - ref[v]["alpha"])[0]
        ok = ~diff_nc
        err_c = max(err_c, float(dc[ok].max(initial=0.0))); err_d = max(err_d, float(dd[ok].max(initial=0.0)))
        err_a = max(err_a, float(da[ok].max(initial=0.0)))
        bad_px += int((((dc > ABS_TOL) | (dd > ABS_TOL) | (da > ABS_TOL)) & ok).sum())
    m["n_contrib_mismatch"] = nc_bad
    m["pixels"] = V * H * W
    m["final_T_maxerr"] = fT_err
    m["color_maxerr"] = err_c; m["depth_maxerr"] = err_d; m["alpha_maxerr"] = err_a
    m["pixels_over_tol"] = bad_px
    if with_backward:
        keep = (~flips).astype(np.float32)[:, None]
        gC = rng.normal(size=(V, 3, H, W)).astype(np.float32) * keep
        gD = rng.normal(size=(V, 1, H, W)).astype(np.float32) * keep
        gA = rng.normal(size=(V, 1, H, W)).astype(np.float32) * keep
        bkey = ("bwd", cache_key, seed, bool(noise_floor), hash(flips.tobytes())) if cache_key is not None else None
        if bkey is not None and bkey in _ORACLE_CACHE:
            ref_g, ref_f32 = _ORACLE_CACHE[bkey]
        else:
            ref_g = oracle_backward(ref, gC, gD, gA)
            ref_f32 = oracle_backward(ref, gC, gD, gA, f32_replay=True) if noise_floor else None
            if bkey is not None:
                _ORACLE_CACHE[bkey] = (ref_g, ref_f32)
        g = cuda_backward(out, gC, gD, gA)
        for k, a in ref_g.items():
            if k == "acc2d":
                continue
            m["grad_relerr_" + k] = grad_rel_err(g[k].cpu().numpy().astype(np.float64).reshape(a.shape), a)
            if ref_f32 is not None:
                m["fp32_floor_" + k] = grad_rel_err(ref_f32[k], a)
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
            # where the all-fp32 reference algorithm itself is further than REL_TOL from the accurate gradient
            # (alpha at the 0.99 cap), the bound is that measured noise floor: we must not be worse than it
            floor = m.get("fp32_floor_" + k[len("grad_relerr_"):], 0.0)
            assert v <= max(REL_TOL, floor), (k, v, m)
