"""Where does a small step (few views per rank, the 8-GPU regime) spend its time?  For each view count: GPU time per
step with / without per-stage events, host enqueue time per step, and the same step replayed from a CUDA graph.
    python tools/probe_step.py [views=3,24] [steps=50]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from topo4d_b200 import engine  # noqa: E402


def main():
    views = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,24").split(",")]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    px_list = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    for nv in views:
        sys_argv = sys.argv
        sys.argv = ["bench.py", "--views", str(nv)]
        a = bench.parse()
        sys.argv = sys_argv
        scene, cams = bench.workload(a)
        H, W = a.height, a.width
        t = {k: torch.from_numpy(v).to(dev) for k, v in scene.items()}
        cam = torch.tensor(engine.pack_cameras_numpy(cams, (0.0, 0.0, 0.0)), device=dev)
        for px in px_list:
            def fwd(ev=None, cap=None):
                return engine.forward(t["means3D"], t["opacities"], cam, H, W, shs=t.get("shs"), scales=t["scales"],
                                      rotations=t["rotations"], sh_degree=a.sh_degree, check="none" if cap else "sync",
                                      cap_instances=cap, stage_events=ev, blend_px=px or None)
            color, radii, depth, alpha, st = fwd()
            s = st.status()
            cap = int(s.num_instances * 1.1) + 4096
            gimg = (torch.sign(color - 0.5) / (3 * H * W), torch.full_like(depth, 0.1 / (H * W)), torch.full_like(alpha, 0.1 / (H * W)))
            del color, depth, alpha, st
            flat = torch.empty(engine.backward(fwd(cap=cap)[-1], *gimg).flat.numel(), dtype=torch.float32, device=dev)

            def step(ev=None):
                *_, st = fwd(ev, cap)
                return engine.backward(st, *gimg, flat=flat, stage_events=ev)

            res = {"views": nv, "blend_px": px, "active_tiles": int(s.num_active_tiles), "I": int(s.num_instances)}
            for mode in ("plain", "events"):
                for _ in range(5):
                    step({} if mode == "events" else None)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev = {} if mode == "events" else None
                h0 = time.perf_counter()
                e0.record()
                for _ in range(steps):
                    step(ev)
                e1.record()
                h1 = time.perf_counter()
                torch.cuda.synchronize()
                res[mode + "_gpu_ms"] = e0.elapsed_time(e1) / steps
                res[mode + "_host_enqueue_ms"] = (h1 - h0) * 1e3 / steps
                if ev:
                    res["stage_ms"] = {k: round(sum(x.elapsed_time(y) for x, y in v) / len(v), 4) for k, v in ev.items()}
            # the same step captured once and replayed as a CUDA graph: no host work, minimal launch gaps
            try:
                g = torch.cuda.CUDAGraph()
                sside = torch.cuda.Stream()
                sside.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(sside):
                    for _ in range(3):
                        step()
                torch.cuda.current_stream().wait_stream(sside)
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    step()
                for _ in range(5):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                res["graph_gpu_ms"] = e0.elapsed_time(e1) / steps
            except Exception as e:  # noqa: BLE001
                res["graph_error"] = repr(e)[:300]
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
