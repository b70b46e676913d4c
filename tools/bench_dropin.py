"""Reference regime through the drop-in modules: one view per Adam step, 8 280 mesh-bound Gaussians, 512x375,
colors_precomp, opacity 1 (train.py:138-146, 661-673, 771) -- what `train.py`'s geometry loop does 7000x per frame.
Wall-clock per WHOLE iteration (render -> image loss -> backward -> optimiser step) in four configurations:
  torch_tail   our rasterizer + the reference's own PyTorch activations (helpers.py:91-112), loss expression
               (train.py:310,317) and torch.optim.Adam
  fused_tail   our rasterizer + fused activations, fused image loss (t4d_image_loss) and FusedAdam, eager
  graph        the same, one CUDA graph per camera (topo4d_b200.graph.capture), replayed
  raster_only  render + L1 + backward only (no SSIM / optimiser): the number earlier rounds reported
    python tools/bench_dropin.py [iters=300] [--json out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizer as Renderer  # noqa: E402
from topo4d_b200 import activations, graph, losses, optim, synth  # noqa: E402

N, W, H, NCAM = 8280, 512, 375, 24
# the reference's parametrisation and learning rates (train.py:120-160, 272-297): raw log-scales / logits / unnormalised
# quaternions, activated by params2rendervar (helpers.py:91-112) on every iteration.  (The r01h numbers in profiles/ were
# taken with an earlier version of this tool that applied these learning rates to ACTIVATED scales, which made the splats
# inflate during the run and the iterations heavier.)
LRS = {"means3D": 0.000016, "rgb_colors": 0.0025, "unnorm_rotations": 0.001, "logit_opacities": 0.0, "log_scales": 0.001,
       "cam_m": 1e-4, "cam_c": 1e-4}


def make_state(dev):
    sc = synth.head_scene(N, seed=0, sh_degree=None, opacity="topo4d")
    cams = synth.ring_cameras(NCAM, w=W, h=H, radius=0.6, focal_over_h=1.6)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    raw = {"means3D": t["means3D"], "rgb_colors": t["colors_precomp"], "unnorm_rotations": t["rotations"],
           "logit_opacities": torch.log(torch.full_like(t["opacities"], 0.9999) / (1 - 0.9999)),       # train.py:142
           "log_scales": torch.log(t["scales"])}
    params = {k: torch.nn.Parameter(v.float().contiguous()) for k, v in raw.items()}
    params["cam_m"] = torch.nn.Parameter(torch.zeros(NCAM, 3, device=dev))
    params["cam_c"] = torch.nn.Parameter(torch.zeros(NCAM, 3, device=dev))
    settings = []
    for c in cams:
        w2c = torch.tensor(c.w2c, dtype=torch.float32, device=dev)
        settings.append(Camera(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                               bg=torch.zeros(3, device=dev), scale_modifier=1.0, viewmatrix=w2c.unsqueeze(0).transpose(1, 2),
                               projmatrix=torch.tensor(c.projmatrix, device=dev).unsqueeze(0), sh_degree=0,
                               campos=torch.tensor(c.campos, device=dev), prefiltered=False, debug=False))
    gts = [torch.rand(3, H, W, device=dev) for _ in range(NCAM)]
    return params, settings, gts


def render(params, cam, fused=False):
    if fused:
        rendervar = activations.params2rendervar(params)
    else:                                                    # helpers.py:91-100 as the reference writes it
        rendervar = {"means3D": params["means3D"], "colors_precomp": params["rgb_colors"],
                     "rotations": torch.nn.functional.normalize(params["unnorm_rotations"]),
                     "opacities": torch.sigmoid(params["logit_opacities"]), "scales": torch.exp(params["log_scales"]),
                     "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}
    return Renderer(raster_settings=cam)(**rendervar)[0]


def torch_loss(im, gt, m, c, win2d):
    x = torch.exp(m)[:, None, None] * im + c[:, None, None]
    conv = lambda t: torch.nn.functional.conv2d(t[None], win2d, padding=5, groups=3)      # noqa: E731
    mu1, mu2 = conv(x), conv(gt)
    s1, s2, s12 = conv(x * x) - mu1 * mu1, conv(gt * gt) - mu2 * mu2, conv(x * gt) - mu1 * mu2
    ss = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    return 0.8 * torch.abs(x - gt).mean() + 0.2 * (1.0 - ss)


def time_loop(fn, iters):
    for i in range(30):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6, e0.elapsed_time(e1) * 1e3 / iters


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300
    dev = "cuda"
    g = torch.tensor([pow(2.718281828459045, -(i - 5) ** 2 / 4.5) for i in range(11)], device=dev)
    g = g / g.sum()
    win2d = (g[:, None] @ g[None, :]).expand(3, 1, 11, 11).contiguous()
    res = {"regime": "reference geometry loop: 1 view/step, 8280 Gaussians, 512x375, colors_precomp, opacity 1", "iters": iters}

    params, settings, gts = make_state(dev)

    def raster_only(i):
        im = render(params, settings[i % NCAM])
        (im - gts[i % NCAM]).abs().mean().backward()
        for p in params.values():
            p.grad = None
    res["raster_only_wall_us"], res["raster_only_device_us"] = time_loop(raster_only, iters)

    params, settings, gts = make_state(dev)
    opt_t = torch.optim.Adam([{"params": [v], "name": k, "lr": LRS[k]} for k, v in params.items()], lr=0.0, eps=1e-15)

    def torch_tail(i):
        k = i % NCAM
        im = render(params, settings[k])
        torch_loss(im, gts[k], params["cam_m"][k], params["cam_c"][k], win2d).backward()
        opt_t.step()
        opt_t.zero_grad(set_to_none=True)
    res["torch_tail_wall_us"], res["torch_tail_device_us"] = time_loop(torch_tail, iters)

    params, settings, gts = make_state(dev)
    opt_f = optim.FusedAdam([{"params": [v], "name": k, "lr": LRS[k]} for k, v in params.items()], lr=0.0, eps=1e-15)

    def fused_tail(i):
        k = i % NCAM
        im = render(params, settings[k], fused=True)
        losses.image_loss(im, gts[k], params["cam_m"][k], params["cam_c"][k]).backward()
        opt_f.step()
        opt_f.zero_grad(set_to_none=True)
    res["fused_tail_wall_us"], res["fused_tail_device_us"] = time_loop(fused_tail, iters)

    params, settings, gts = make_state(dev)
    opt_g = optim.FusedAdam([{"params": [v], "name": k, "lr": LRS[k]} for k, v in params.items()], lr=0.0, eps=1e-15, capturable=True)

    def make_iter(k):
        def it():
            im = render(params, settings[k], fused=True)
            loss = losses.image_loss(im, gts[k], params["cam_m"][k], params["cam_c"][k])
            loss.backward()
            opt_g.step()
            opt_g.zero_grad(set_to_none=True)
            return loss
        return it
    t0 = time.perf_counter()
    steps = [graph.capture(make_iter(k), capacity_headroom=4.0) for k in range(NCAM)]    # random targets inflate the splats
    res["graph_capture_s_for_24_cameras"] = time.perf_counter() - t0
    res["graph_wall_us"], res["graph_device_us"] = time_loop(lambda i: steps[i % NCAM].replay(), iters)
    for s in steps:
        s.check()
    res["graph_final_loss"] = float(steps[0].outputs.item())
    res["mpix_per_s_graph"] = W * H / res["graph_wall_us"]
    print(json.dumps(res))
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
