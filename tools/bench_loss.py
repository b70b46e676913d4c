"""Fused image loss (fwd+bwd, t4d_image_loss) and fused Adam against the reference's PyTorch formulation on the same GPU.
    python tools/bench_loss.py [--json out.json]
Times with CUDA events after warm-up; the PyTorch arm is the reference's own expression (train.py:310,317 with
external.py:85-116 / helpers.py:115-116) run through autograd in fp32 on the device."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from topo4d_b200 import losses, optim  # noqa: E402

DEV = "cuda:0"


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def torch_loss(win2d):
    def conv(t):
        return torch.nn.functional.conv2d(t, win2d, padding=5, groups=3)

    def f(render, target, m, c):
        im = torch.exp(m)[:, :, None, None] * render + c[:, :, None, None]
        l1 = torch.abs(im - target).mean(dim=(1, 2, 3))
        mu1, mu2 = conv(im), conv(target)
        s1, s2, s12 = conv(im * im) - mu1 * mu1, conv(target * target) - mu2 * mu2, conv(im * target) - mu1 * mu2
        ss = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean(dim=(1, 2, 3))
        return (0.8 * l1 + 0.2 * (1 - ss)).sum()
    return f


def main():
    g = torch.tensor([pow(2.718281828459045, -(i - 5) ** 2 / 4.5) for i in range(11)], device=DEV)
    g = g / g.sum()
    win2d = (g[:, None] @ g[None, :]).expand(3, 1, 11, 11).contiguous()
    ref = torch_loss(win2d)
    rows = []
    for v, h, w in ((1, 375, 512), (1, 1080, 1920), (24, 1080, 1920)):
        render = torch.rand(v, 3, h, w, device=DEV, requires_grad=True)
        target = torch.rand(v, 3, h, w, device=DEV)
        m = torch.zeros(v, 3, device=DEV, requires_grad=True)
        c = torch.zeros(v, 3, device=DEV, requires_grad=True)

        def ours():
            render.grad = None
            losses.image_loss(render, target, m, c).backward()

        def theirs():
            render.grad = None
            ref(render, target, m, c).backward()
        t_ours = timed(ours)
        t_ref = timed(theirs, iters=10, warm=3)
        px = v * h * w
        # algorithmic bytes: read render + target, write dL/drender (3 channels fp32 each)
        rows.append({"views": v, "h": h, "w": w, "fused_ms": t_ours, "pytorch_ms": t_ref, "speedup": t_ref / t_ours,
                     "fused_Mpix_per_s": px / 1e6 / (t_ours * 1e-3), "algorithmic_GBps": px * 36 / 1e9 / (t_ours * 1e-3)})
        print(json.dumps(rows[-1]), flush=True)
        del render, target
    # Adam over the SH-3 parameter set of config 2 (60k Gaussians: 3 + 48 + 4 + 1 + 3 floats each)
    shapes = [(60000, 3), (60000, 48), (60000, 4), (60000, 1), (60000, 3), (24, 3), (24, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    for p in ps:
        p.grad = torch.randn_like(p)
    fused = optim.FusedAdam([{"params": [p], "lr": 1e-3} for p in ps], lr=0.0, eps=1e-15)
    stock = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in ps], lr=0.0, eps=1e-15)
    n = sum(p.numel() for p in ps)
    tf, ts = timed(fused.step), timed(stock.step)
    rows.append({"adam_elements": n, "fused_ms": tf, "torch_adam_ms": ts, "speedup": ts / tf, "fused_GBps": n * 28 / 1e9 / (tf * 1e-3)})
    print(json.dumps(rows[-1]), flush=True)
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
