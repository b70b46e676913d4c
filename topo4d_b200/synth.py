"""Seeded synthetic scenes and cameras for tests, smoke() and bench.py (numpy only).

Workloads are the ones SURVEY.md section 8(d) / BASELINE.json name:
  * config 1: 1 view 256x256, 5k random Gaussians  -> :func:`random_scene`
  * config 2/3: 24 views 1920x1080, ~60k mesh-bound Gaussians on a head-sized ellipsoid,
    SH degree 3 -> :func:`head_scene` + :func:`ring_cameras`
  * config 4: 8192^2 UV-atlas stand-in mesh for the face3d bake -> :func:`uv_grid_mesh`

Camera construction follows the reference's ``setup_camera`` (helpers.py:63-88): the op
receives ``viewmatrix = w2c^T`` and ``projmatrix = (P @ w2c)^T`` as 16 consecutive floats,
i.e. element ``[4*col+row]`` of the mathematical matrix, with the z in [0,1] projection
and near/far = 0.01/100 (train.py:98).  Unlike the reference (helpers.py:66 takes row 3 of
the inverse, always zero) ``campos`` here is the true camera centre, which SH needs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Camera:
    """Plain-number twin of ``GaussianRasterizationSettings`` (no torch, no device)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] float32 = w2c^T
    projmatrix: np.ndarray   # [4,4] float32 = (P @ w2c)^T
    campos: np.ndarray       # [3] float32
    w2c: np.ndarray          # [4,4] float64


def make_camera(w2c: np.ndarray, w: int, h: int, fx: float, fy: float, cx: float | None = None,
                cy: float | None = None, near: float = 0.01, far: float = 100.0) -> Camera:
    """K, w2c -> op matrices, same formulas as helpers.py:63-88 (float32 like the reference)."""
    cx = w / 2.0 if cx is None else cx
    cy = h / 2.0 if cy is None else cy
    w2c32 = np.asarray(w2c, np.float32)
    view = w2c32.T.copy()
    proj = np.array([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                     [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                     [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                     [0.0, 0.0, 1.0, 0.0]], np.float32)
    full = (view @ proj.T).astype(np.float32)
    campos = np.linalg.inv(np.asarray(w2c, np.float64))[:3, 3].astype(np.float32)
    return Camera(int(h), int(w), w / (2 * fx), h / (2 * fy), view, full, campos, np.asarray(w2c, np.float64))


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """OpenCV-style camera (x right, y down, z forward) world-to-camera 4x4."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(target, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f])
    w2c = np.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3] = -R @ eye
    return w2c


def ring_cameras(n_views: int = 24, w: int = 1920, h: int = 1080, radius: float = 1.0,
                 elev_deg: float = 15.0, focal_over_h: float = 1.6) -> list[Camera]:
    """Two rings (elevation +-elev) looking at the origin; fy = 1.6 H (head ~55 % of height)."""
    cams = []
    per = (n_views + 1) // 2
    for i in range(n_views):
        ring, k = divmod(i, per)
        az = 2 * math.pi * (k + 0.5 * ring) / per
        el = math.radians(elev_deg if ring == 0 else -elev_deg)
        eye = radius * np.array([math.cos(el) * math.sin(az), math.sin(el), -math.cos(el) * math.cos(az)])
        f = focal_over_h * h
        cams.append(make_camera(look_at(eye), w, h, f, f))
    return cams


def front_camera(w: int = 256, h: int = 256, dist: float = 4.0, fx: float | None = None) -> Camera:
    """Config-1 camera: at (0,0,-dist) looking +z, fx = fy = w (tanfov 0.5)."""
    fx = float(w) if fx is None else fx
    return make_camera(look_at((0.0, 0.0, -dist), (0.0, 0.0, 0.0), (0.0, -1.0, 0.0)), w, h, fx, fx)


def random_scene(n: int = 5000, seed: int = 0, sh_degree: int | None = None, extent: float = 1.0):
    """Config 1: means ~U([-1,1]^3), anisotropic log-uniform scales in [0.01,0.1], random unit
    quaternions, opacity ~U(0.05,1), colours ~U(0,1) (or SH coefficients when sh_degree given)."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-extent, extent, (n, 3))
    scales = np.exp(rng.uniform(math.log(0.01), math.log(0.1), (n, 3)))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = rng.uniform(0.05, 1.0, (n, 1))
    out = dict(means3D=means, scales=scales, rotations=q, opacities=opac)
    if sh_degree is None:
        out["colors_precomp"] = rng.uniform(0, 1, (n, 3))
    else:
        k = (sh_degree + 1) ** 2
        sh = rng.normal(0, 0.3, (n, k, 3))
        sh[:, 0, :] = (rng.uniform(0, 1, (n, 3)) - 0.5) / 0.28209479177387814
        out["shs"] = sh
    return {k: np.ascontiguousarray(v, np.float32) for k, v in out.items()}


def _quat_from_normals(nrm: np.ndarray) -> np.ndarray:
    """Same construction as the reference's build_quaterion (external.py:45-61) followed by the
    normalisation params2rendervar applies (helpers.py:95)."""
    u = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    axis0 = np.array([1.0, 0.0, 0.0])
    ax = np.cross(np.broadcast_to(axis0, u.shape), u)
    ang = np.arccos(np.clip(u @ axis0, -1.0, 1.0))
    q = np.concatenate([np.cos(ang / 2)[:, None], ax * np.sin(ang / 2)[:, None]], axis=1)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def head_scene(n: int = 60000, seed: int = 0, sh_degree: int | None = 3, opacity: str = "topo4d"):
    """Config 2: ~n mesh-bound Gaussians on a head-like ellipsoid (semi-axes 0.09x0.12x0.10 m,
    +-5 % low-frequency radial noise); isotropic scale = half nearest-neighbour distance
    (train.py:132-143) x exp(N(0,0.1)); rotation from the vertex normal (train.py:136);
    opacity 'topo4d' = 1.0 (sigmoid(1000), train.py:142) or 'generic' = U(0.3,1)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = math.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], 1)
    # low-frequency radial noise
    amp = rng.normal(size=(6,)) * 0.02
    frq = rng.integers(1, 4, size=(6, 3)).astype(np.float64)
    pha = rng.uniform(0, 2 * math.pi, size=(6,))
    noise = sum(amp[k] * np.sin(d @ frq[k] + pha[k]) for k in range(6))
    noise = np.clip(noise, -0.05, 0.05)
    axes = np.array([0.09, 0.12, 0.10])
    pts = d * axes * (1.0 + noise)[:, None]
    nrm = d / axes
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dist, _ = cKDTree(pts).query(pts, k=2)
    sc = 0.5 * np.maximum(dist[:, 1], 1e-7) * np.exp(rng.normal(0, 0.1, n))
    scales = np.repeat(sc[:, None], 3, 1)
    q = _quat_from_normals(nrm)
    if opacity == "topo4d":
        opac = np.ones((n, 1))
    else:
        opac = rng.uniform(0.3, 1.0, (n, 1))
    albedo = 0.5 + 0.3 * np.stack([np.sin(7 * d[:, 0] + 1.0), np.sin(5 * d[:, 1] + 2.0), np.sin(6 * d[:, 2] + 3.0)], 1)
    out = dict(means3D=pts, scales=scales, rotations=q, opacities=opac)
    if sh_degree is None:
        out["colors_precomp"] = albedo
    else:
        k = (sh_degree + 1) ** 2
        sh = rng.normal(0, 0.05, (n, k, 3))
        sh[:, 0, :] = (albedo - 0.5) / 0.28209479177387814
        out["shs"] = sh
    return {k: np.ascontiguousarray(v, np.float32) for k, v in out.items()}


def uv_grid_mesh(grid: int = 245, res: int = 8192, seed: int = 0, jitter: float = 0.3, extras: bool = True):
    """Config 4: regular grid x grid quad mesh -> 2*grid^2 triangles over [0,res-1]^2, z = 0
    (process_uv, helpers.py:945-950), per-vertex jitter <= `jitter` cell, colours U(0,1);
    `extras` appends a few degenerate and border-crossing triangles."""
    rng = np.random.default_rng(seed)
    g1 = grid + 1
    cell = (res - 1) / grid
    ys, xs = np.meshgrid(np.arange(g1), np.arange(g1), indexing="ij")
    v = np.stack([xs * cell, ys * cell], -1).reshape(-1, 2).astype(np.float64)
    interior = ((xs > 0) & (xs < grid) & (ys > 0) & (ys < grid)).reshape(-1)
    v[interior] += rng.uniform(-jitter, jitter, (int(interior.sum()), 2)) * cell
    idx = (ys * g1 + xs)[:-1, :-1].reshape(-1)
    t0 = np.stack([idx, idx + 1, idx + g1], 1)
    t1 = np.stack([idx + 1, idx + g1 + 1, idx + g1], 1)
    tris = np.concatenate([t0, t1], 0)
    verts = np.concatenate([v, np.zeros((v.shape[0], 1))], 1)
    if extras:
        nv = verts.shape[0]
        ex = np.array([[10.0, 10.0, 0], [10.0, 10.0, 0], [10.0, 10.0, 0],            # point-degenerate
                       [50.5, 60.5, 0], [80.5, 60.5, 0], [110.5, 60.5, 0],           # collinear
                       [-20.0, -30.0, 0], [40.0, -10.0, 0], [-5.0, 35.0, 0],         # crosses the border
                       [res + 10.0, res - 30.0, 0], [res - 40.0, res + 5.0, 0], [res - 25.0, res - 45.0, 0]])
        verts = np.concatenate([verts, ex], 0)
        tris = np.concatenate([tris, nv + np.arange(12).reshape(4, 3)], 0)
    colors = rng.uniform(0, 1, (verts.shape[0], 3))
    return verts.astype(np.float64), tris.astype(np.int64), colors.astype(np.float64)


def dense_head_scene(n: int = 4_000_000, seed: int = 0, opacity: float = 1.0):
    """Texture-stage stand-in (4-5 M UV-densified Gaussians, colours optimised against ~4096x3000 photos,
    helpers.py:608-609, train.py:715-743): n points on the same head ellipsoid as :func:`head_scene`, isotropic scale = half
    the local point spacing (analytic: sqrt(surface area per point) -- a k-d tree over millions of points would dominate the
    set-up), rotation from the normal, ``colors_precomp`` colours (the reference never uses SH, helpers.py:82,94,105)."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = math.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], 1)
    axes = np.array([0.09, 0.12, 0.10])
    pts = d * axes
    nrm = d / axes
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    # area element of the ellipsoid per unit-sphere area: |axes-scaled normal| factor
    stretch = np.linalg.norm(d / axes, axis=1) * axes.prod()
    spacing = np.sqrt(4 * math.pi * stretch / n)
    sc = 0.5 * spacing * np.exp(rng.normal(0, 0.1, n))
    albedo = 0.5 + 0.3 * np.stack([np.sin(70 * d[:, 0] + 1.0), np.sin(50 * d[:, 1] + 2.0), np.sin(60 * d[:, 2] + 3.0)], 1)
    out = dict(means3D=pts, scales=np.repeat(sc[:, None], 3, 1), rotations=_quat_from_normals(nrm),
               opacities=np.full((n, 1), opacity), colors_precomp=albedo)
    return {k: np.ascontiguousarray(v, np.float32) for k, v in out.items()}
