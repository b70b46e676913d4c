"""One camera of the reference regime: capture the whole iteration, replay it a few times (for ncu / timing)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools.bench_dropin as bd  # noqa: E402
from topo4d_b200 import graph, losses, optim  # noqa: E402

dev = "cuda"
params, settings, gts = bd.make_state(dev)
opt = optim.FusedAdam([{"params": [v], "name": k, "lr": bd.LRS[k]} for k, v in params.items()], lr=0.0, eps=1e-15, capturable=True)


def make(k):
    def it():
        im = bd.render(params, settings[k])
        loss = losses.image_loss(im, gts[k], params["cam_m"][k], params["cam_c"][k])
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss
    return it


n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ncam = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = [graph.capture(make(k), capacity_headroom=4.0) for k in range(ncam)]
for rounds in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        steps[i % ncam].replay()
    e1.record()
    torch.cuda.synchronize()
    print("cameras", ncam, "replay us/iter:", e0.elapsed_time(e1) * 1e3 / n, "wall", (time.perf_counter() - t0) / n * 1e6,
          "instances", int(steps[0].states[0].status().num_instances), flush=True)
