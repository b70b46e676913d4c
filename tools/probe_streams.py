"""Does running the views of a step as independent pipelines on separate streams (captured in one CUDA graph) beat one
batched launch sequence?  Every kernel of a launch ends in a serial tail (the longest tile list is walked by one warp while
the rest of the GPU drains); pipelines on different streams fill each other's tails.
    python tools/probe_streams.py VIEWS GROUPSIZE[,GROUPSIZE...] [steps=30] [px=0]
prints one JSON line per group size: graph-replayed ms per step (all VIEWS views forward + backward, gradients summed)."""
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
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [nv, 1]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    px = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
    ref = None
    for gsize in sizes:
        # interleaved ownership (group i takes views i, i + n, ...), like the ranks of the view-parallel step
        ngr = (nv + gsize - 1) // gsize
        groups = [list(range(i, nv, ngr)) for i in range(ngr)]
        cams_g = [cam[g].contiguous() for g in groups]
        caps, gimgs, flats = [], [], []
        for cg in cams_g:
            for _ in range(2):      # the second pass picks the blend variant from the first one's tile count
                color, radii, depth, alpha, st = engine.forward(t["means3D"], t["opacities"], cg, H, W, shs=t["shs"], scales=t["scales"],
                                                                rotations=t["rotations"], sh_degree=a.sh_degree, blend_px=px or None)
                st.status()
            caps.append(int(st.status().num_instances * 1.1) + 4096)
            gimgs.append((torch.sign(color - 0.5) / (3 * H * W), torch.full_like(depth, 0.1 / (H * W)), torch.full_like(alpha, 0.1 / (H * W))))
            flats.append(torch.empty_like(engine.backward(st, *gimgs[-1]).flat))
            del color, depth, alpha, st
        streams = [torch.cuda.Stream() for _ in groups]

        def step():
            cur = torch.cuda.current_stream()
            for i, cg in enumerate(cams_g):
                s = streams[i] if len(groups) > 1 else cur
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    *_, st = engine.forward(t["means3D"], t["opacities"], cg, H, W, shs=t["shs"], scales=t["scales"],
                                            rotations=t["rotations"], sh_degree=a.sh_degree, check="none", cap_instances=caps[i],
                                            blend_px=px or None)
                    engine.backward(st, *gimgs[i], flat=flats[i])
            for i in range(len(groups)):
                if len(groups) > 1:
                    cur.wait_stream(streams[i])
            for f in flats[1:]:
                flats[0].add_(f)
            return flats[0]

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = step()
        ms = timed_replay(g, steps)
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        print(json.dumps({"views": nv, "views_per_pipeline": gsize, "pipelines": len(groups), "blend_px": px or "auto",
                          "graph_ms_per_step": round(ms, 4), "grad_maxdiff_vs_first": err}), flush=True)
        del g, out, flats, gimgs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
