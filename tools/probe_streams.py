"""Does running the views of a small step as independent pipelines on separate streams (captured in one CUDA graph)
beat one batched launch sequence?  python tools/probe_streams.py [views=3] [steps=50]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from topo4d_b200 import engine  # noqa: E402


def timed_replay(g, steps):
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    argv = sys.argv
    sys.argv = ["bench.py", "--views", str(nv)]
    a = bench.parse()
    sys.argv = argv
    scene, cams = bench.workload(a)
    H, W = a.height, a.width
    t = {k: torch.from_numpy(v).to(dev) for k, v in scene.items()}
    cam = torch.tensor(engine.pack_cameras_numpy(cams, (0.0, 0.0, 0.0)), device=dev)
    for gsize in (nv, 1):
        for px in (1, 2):
            groups = [list(range(i, min(i + gsize, nv))) for i in range(0, nv, gsize)]
            cams_g = [cam[g].contiguous() for g in groups]
            caps, gimgs, flats = [], [], []
            for cg in cams_g:
                color, radii, depth, alpha, st = engine.forward(t["means3D"], t["opacities"], cg, H, W, shs=t["shs"], scales=t["scales"],
                                                                rotations=t["rotations"], sh_degree=a.sh_degree, blend_px=px)
                caps.append(int(st.status().num_instances * 1.1) + 4096)
                gimgs.append((torch.sign(color - 0.5) / (3 * H * W), torch.full_like(depth, 0.1 / (H * W)), torch.full_like(alpha, 0.1 / (H * W))))
                flats.append(torch.empty_like(engine.backward(st, *gimgs[-1]).flat))
            streams = [torch.cuda.Stream() for _ in groups]

            def step():
                cur = torch.cuda.current_stream()
                for i, cg in enumerate(cams_g):
                    s = streams[i] if len(groups) > 1 else cur
                    s.wait_stream(cur)
                    with torch.cuda.stream(s):
                        *_, st = engine.forward(t["means3D"], t["opacities"], cg, H, W, shs=t["shs"], scales=t["scales"],
                                                rotations=t["rotations"], sh_degree=a.sh_degree, check="none", cap_instances=caps[i],
                                                blend_px=px)
                        engine.backward(st, *gimgs[i], flat=flats[i])
                for i in range(len(groups)):
                    if len(groups) > 1:
                        cur.wait_stream(streams[i])
                if len(groups) > 1:
                    for f in flats[1:]:
                        flats[0].add_(f)

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            print(json.dumps({"views": nv, "views_per_pipeline": gsize, "pipelines": len(groups), "blend_px": px,
                              "graph_ms_per_step": timed_replay(g, steps)}), flush=True)


if __name__ == "__main__":
    main()
