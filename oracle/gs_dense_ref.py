"""Dense fp64 autograd formulation of the Gaussian rasterizer (TEST INFRASTRUCTURE).

An independent second statement of SURVEY.md Appendix A used to validate
oracle/gs_oracle.c: no tiles-as-loops, no hand-written backward -- every pixel sees every
Gaussian in global depth order, masked by the Gaussian's 16x16-tile rectangle, and the
gradients come from ``torch.autograd`` with the op's differentiability conventions (A.7):
straight-through 0.99 alpha cap, hard masks carry no gradient, frustum-clamped t.x/t.y are
constants.  Small scenes only (memory is pixels x Gaussians).  PARITY UNPINNED, like the C
oracle: the reference's rasterizer source is not vendored.

Conventions follow the reference call sites: helpers.py:63-88 (matrices), helpers.py:91-112
(inputs), external.py:26-43 (quaternion -> R), helpers.py:836-922 (SH).
"""
import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def _sh_color(deg, sh, dirs):
    """sh [N,K,3], dirs [N,3] unit -> [N,3] (before +0.5 / clamp).  Order as helpers.py:867-922."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def render_dense(means3D, means2D, opacities, *, image_height, image_width, tanfovx, tanfovy, bg, viewmatrix,
                 projmatrix, campos, sh_degree=0, scale_modifier=1.0, shs=None, colors_precomp=None,
                 scales=None, rotations=None, cov3D_precomp=None, rect=None, dtype=torch.float64):
    """Returns (color[3,H,W], radii[N], depth[1,H,W], alpha[1,H,W]).  All tensor inputs may require grad.
    `rect` ([N,4] int tile rectangle, max exclusive) overrides the internally computed one so that
    fp64-vs-fp32 rounding of the (non-differentiable) radius cannot change the instance lists."""
    H, W = int(image_height), int(image_width)
    t = lambda a: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=dtype)
    V = t(viewmatrix).reshape(16)
    P = t(projmatrix).reshape(16)
    cam = t(campos).reshape(3)
    bgc = t(bg).reshape(3)
    N = means3D.shape[0]
    p = means3D.to(dtype)
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    tx = V[0] * px + V[4] * py + V[8] * pz + V[12]
    ty = V[1] * px + V[5] * py + V[9] * pz + V[13]
    tz = V[2] * px + V[6] * py + V[10] * pz + V[14]
    hx = P[0] * px + P[4] * py + P[8] * pz + P[12]
    hy = P[1] * px + P[5] * py + P[9] * pz + P[13]
    hw = P[3] * px + P[7] * py + P[11] * pz + P[15]
    pw = 1.0 / (hw + 0.0000001)
    ndc = torch.stack([hx * pw, hy * pw], 1) + means2D.to(dtype)[:, :2]      # means2D: NDC-unit offset (API contract)
    vis = tz.detach() > 0.2

    if cov3D_precomp is not None:
        c = cov3D_precomp.to(dtype)
        Sig = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], 1), torch.stack([c[:, 1], c[:, 3], c[:, 4]], 1),
                           torch.stack([c[:, 2], c[:, 4], c[:, 5]], 1)], 1)
    else:
        s = scale_modifier * scales.to(dtype)
        q = rotations.to(dtype)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
                         torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
                         torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
        A = R * s[:, None, :]
        Sig = A @ A.transpose(1, 2)

    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tzs = torch.where(vis, tz, torch.ones_like(tz))
    txtz, tytz = tx / tzs, ty / tzs
    inx = (txtz.detach() >= -limx) & (txtz.detach() <= limx)
    iny = (tytz.detach() >= -limy) & (tytz.detach() <= limy)
    cx = torch.where(inx, tx, (txtz.clamp(-limx, limx) * tzs).detach())
    cy = torch.where(iny, ty, (tytz.clamp(-limy, limy) * tzs).detach())
    zero = torch.zeros_like(tzs)
    J = torch.stack([torch.stack([fx / tzs, zero, -(fx * cx) / (tzs * tzs)], 1),
                     torch.stack([zero, fy / tzs, -(fy * cy) / (tzs * tzs)], 1)], 1)        # [N,2,3]
    Wr = torch.stack([torch.stack([V[0], V[4], V[8]]), torch.stack([V[1], V[5], V[9]]), torch.stack([V[2], V[6], V[10]])])
    T = J @ Wr
    cov = T @ Sig @ T.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = vis & (det.detach() != 0)
    dets = torch.where(ok, det, torch.ones_like(det))
    cA, cB, cC = c / dets, -b / dets, a / dets
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    if rect is None:
        pd = pix.detach()
        rmin_x = torch.clamp(torch.trunc((pd[:, 0] - radius) / 16), 0, gx)
        rmin_y = torch.clamp(torch.trunc((pd[:, 1] - radius) / 16), 0, gy)
        rmax_x = torch.clamp(torch.trunc((pd[:, 0] + radius + 15) / 16), 0, gx)
        rmax_y = torch.clamp(torch.trunc((pd[:, 1] + radius + 15) / 16), 0, gy)
        rect_t = torch.stack([rmin_x, rmin_y, rmax_x, rmax_y], 1).long()
    else:
        rect_t = torch.as_tensor(np.asarray(rect)).long()
    area = (rect_t[:, 2] - rect_t[:, 0]) * (rect_t[:, 3] - rect_t[:, 1])
    ok = ok & (area > 0)
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)

    if shs is not None:
        d = p - cam
        dirs = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp(_sh_color(sh_degree, shs.to(dtype), dirs) + 0.5, min=0.0)
    else:
        rgb = colors_precomp.to(dtype)

    # global order = (float32 depth bits, index), the order a stable (tile|depth) sort gives inside every tile
    order = torch.argsort(tz.detach().to(torch.float32), stable=True)
    order = order[ok[order]]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxl = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).to(dtype)                  # [P,2]
    tix, tiy = (xs.reshape(-1) // 16), (ys.reshape(-1) // 16)
    xy = pix[order]                                                                   # [K,2]
    dx = xy[None, :, 0] - pxl[:, None, 0]
    dy = xy[None, :, 1] - pxl[:, None, 1]
    power = -0.5 * (cA[order][None] * dx * dx + cC[order][None] * dy * dy) - cB[order][None] * dx * dy
    G = torch.exp(torch.clamp(power, max=0.0))
    op = opacities.to(dtype).reshape(-1)[order]
    alpha_raw = op[None] * G
    alpha = alpha_raw + (torch.clamp(alpha_raw, max=0.99) - alpha_raw).detach()     # straight-through cap
    rct = rect_t[order]
    inrect = (tix[:, None] >= rct[None, :, 0]) & (tix[:, None] < rct[None, :, 2]) & \
             (tiy[:, None] >= rct[None, :, 1]) & (tiy[:, None] < rct[None, :, 3])
    valid = inrect & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
    with torch.no_grad():
        # termination decision replays the float32 recurrence test_T = T*(1-alpha) >= 1e-4
        testT = torch.cumprod(1.0 - a_eff, dim=1)
        keep = testT >= 0.0001
    a_fin = a_eff * keep
    one_m = 1.0 - a_fin
    Tincl = torch.cumprod(one_m, dim=1)
    Tbefore = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    wgt = a_fin * Tbefore
    Tfinal = Tincl[:, -1] if Tincl.shape[1] > 0 else torch.ones(H * W, dtype=dtype)
    col = wgt @ rgb[order] + Tfinal[:, None] * bgc[None]
    dep = wgt @ tz[order]
    alp = wgt.sum(1)
    if N == 0:
        col = torch.zeros(H * W, 3, dtype=dtype)
    return col.t().reshape(3, H, W), radii, dep.reshape(1, H, W), alp.reshape(1, H, W)
