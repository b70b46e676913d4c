"""Diff this repository's rasterizer against the REAL upstream extension, for a user who has it.

The reference's rasterizer (`ashawkey/diff-gaussian-rasterization`, README.md:22-24) is not vendored, not installed and cannot be
fetched in the build container, so the oracle restates its published algorithm and a handful of details are ASSUMED (listed in
INTEGRATION.md section 4 -- each marked "(?)" in SURVEY.md Appendix A).  On a machine where `diff_gaussian_rasterization._C` IS
importable (the upstream CUDA build), this script renders the same seeded scenes through both and prints, per assumed
behaviour, the quantity that would reveal a mismatch:

    PYTHONPATH=/path/to/upstream/site-packages python tools/compare_with_upstream.py

It exits 0 and says so when the upstream extension is absent (always the case in the graft containers)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_upstream():
    """The upstream package has the same module name as our shim: import it from a path that does not contain this repo."""
    saved = list(sys.path)
    try:
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
        for m in [k for k in sys.modules if k.startswith("diff_gaussian_rasterization")]:
            del sys.modules[m]
        mod = importlib.import_module("diff_gaussian_rasterization")
        if not hasattr(mod, "_C"):
            return None
        return mod
    except Exception:  # noqa: BLE001
        return None
    finally:
        sys.path = saved
        for m in [k for k in sys.modules if k.startswith("diff_gaussian_rasterization")]:
            del sys.modules[m]


def main():
    up = load_upstream()
    if up is None:
        print("upstream diff_gaussian_rasterization._C is not importable here: nothing to compare (see INTEGRATION.md section 4 for the "
              "assumed behaviours this script would check)")
        return 0
    import torch
    sys.path.insert(0, ROOT)
    from topo4d_b200 import rasterizer as ours, synth

    dev = torch.device("cuda:0")
    worst = {}

    def settings(mod, cam, bg, deg):
        w2c = torch.tensor(cam.w2c, dtype=torch.float32, device=dev)
        return mod.GaussianRasterizationSettings(
            image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
            bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=1.0, viewmatrix=w2c.unsqueeze(0).transpose(1, 2),
            projmatrix=torch.tensor(cam.projmatrix, device=dev).unsqueeze(0), sh_degree=deg, campos=torch.tensor(cam.campos, device=dev),
            prefiltered=False, debug=False)

    def run(mod, scene, cam, bg, deg, gC, gD, gA):
        p = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in scene.items()}
        m2d = torch.zeros_like(p["means3D"], requires_grad=True)
        kw = dict(means3D=p["means3D"], means2D=m2d, opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"])
        kw["shs" if "shs" in p else "colors_precomp"] = p.get("shs", p.get("colors_precomp"))
        color, radii, depth, alpha = mod.GaussianRasterizer(raster_settings=settings(mod, cam, bg, deg))(**kw)
        (color * gC).sum().add((depth * gD).sum()).add((alpha * gA).sum()).backward()
        g = {k: v.grad.detach().cpu().numpy() for k, v in p.items() if v.grad is not None}
        g["means2D"] = m2d.grad.detach().cpu().numpy()
        return dict(color=color.detach().cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.detach().cpu().numpy(),
                    alpha=alpha.detach().cpu().numpy()), g

    cases = [("config-1 precomp", synth.random_scene(5000, seed=0), synth.front_camera(256, 256), (0.2, 0.5, 0.8), 0),
             ("config-1 SH-3", synth.random_scene(5000, seed=13, sh_degree=3), synth.make_camera(synth.look_at((1.5, 0.8, -3.5)), 256, 256, 256.0, 256.0), (0.1, 0.1, 0.1), 3),
             ("topo4d regime", synth.head_scene(8280, seed=0, sh_degree=None, opacity="topo4d"), synth.ring_cameras(2, w=512, h=375, radius=0.6)[0], (0, 0, 0), 0)]
    for name, scene, cam, bg, deg in cases:
        H, W = cam.image_height, cam.image_width
        gen = torch.Generator(device=dev).manual_seed(1)
        gC, gD, gA = (torch.randn(s, device=dev, generator=gen) for s in ((3, H, W), (1, H, W), (1, H, W)))
        a_out, a_g = run(up, scene, cam, bg, deg, gC, gD, gA)
        b_out, b_g = run(ours, scene, cam, bg, deg, gC, gD, gA)
        print(f"== {name}")
        print("  radii mismatches                  :", int((a_out["radii"] != b_out["radii"]).sum()), " (tile rectangles / culling, A.3)")
        for k, what in (("color", "A.5 blend, background term"), ("depth", "A.5: depth un-normalised, no background"),
                        ("alpha", "A.5 (?): alpha = sum alpha_i T_i")):
            e = float(np.abs(a_out[k] - b_out[k]).max())
            worst[k] = max(worst.get(k, 0.0), e)
            print(f"  max |{k:5s} difference|            : {e:.3e}   ({what})")
        for k in sorted(a_g):
            ref = a_g[k].astype(np.float64)
            e = float((np.abs(b_g[k].reshape(ref.shape) - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max() + 1e-30)).max())
            worst["grad_" + k] = max(worst.get("grad_" + k, 0.0), e)
            hint = {"means2D": "A.6/A.7: NDC scaling x(0.5 W, 0.5 H), z = 0", "means3D": "A.7 (?): d depth / d mean = third row of W; clamp masks",
                    "opacities": "A.6: no clamp mask on the 0.99 cap (straight-through)"}.get(k, "")
            print(f"  grad {k:14s} rel. difference : {e:.3e}   {hint}")
    ok = all(worst.get(k, 0) <= 1e-4 for k in ("color", "depth", "alpha")) and all(v <= 1e-3 for k, v in worst.items() if k.startswith("grad_"))
    print("VERDICT:", "within the north-star tolerances (1e-4 abs on images, 1e-3 rel on gradients)" if ok else "DIFFERENCES ABOVE TOLERANCE -- see the hints")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
