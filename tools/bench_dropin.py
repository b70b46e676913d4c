"""Reference regime through the drop-in module: one view per call, 8 280 mesh-bound Gaussians, 512x375,
colors_precomp, opacity 1 (train.py:138-146, 771) -- what `train.py`'s geometry loop does 7000x per frame.
Reports wall-clock per iteration of  Renderer(cam)(**rendervar); loss.backward()  and the device time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizer as Renderer  # noqa: E402
from topo4d_b200 import synth  # noqa: E402


def main():
    n, w, h = 8280, 512, 375
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    sc = synth.head_scene(n, seed=0, sh_degree=None, opacity="topo4d")
    cams = synth.ring_cameras(24, w=w, h=h, radius=0.6, focal_over_h=1.6)
    dev = "cuda"
    params = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in sc.items()}
    settings = []
    for c in cams:
        w2c = torch.tensor(c.w2c, dtype=torch.float32, device=dev)
        settings.append(Camera(image_height=h, image_width=w, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                               bg=torch.zeros(3, device=dev), scale_modifier=1.0, viewmatrix=w2c.unsqueeze(0).transpose(1, 2),
                               projmatrix=torch.tensor(c.projmatrix, device=dev).unsqueeze(0), sh_degree=0,
                               campos=torch.tensor(c.campos, device=dev), prefiltered=False, debug=False))
    target = torch.rand(3, h, w, device=dev)

    def it(i):
        rendervar = {"means3D": params["means3D"], "colors_precomp": params["colors_precomp"],
                     "rotations": torch.nn.functional.normalize(params["rotations"]), "opacities": params["opacities"],
                     "scales": params["scales"], "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}
        im, radius, _, _ = Renderer(raster_settings=settings[i % 24])(**rendervar)
        loss = (im - target).abs().mean()
        loss.backward()
        for p in params.values():
            p.grad = None

    for i in range(30):
        it(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        it(i)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters
    print(json.dumps({"regime": "reference geometry loop: 1 view/call, 8280 Gaussians, 512x375, colors_precomp, opacity 1",
                      "iters": iters, "wall_us_per_iter": wall * 1e6, "device_us_per_iter": e0.elapsed_time(e1) * 1e3 / iters,
                      "mpix_per_s": w * h / 1e6 / wall}))


if __name__ == "__main__":
    main()
