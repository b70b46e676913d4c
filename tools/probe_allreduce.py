"""What does the gradient exchange cost inside the view-parallel step?  Under torchrun: the same local fwd+bwd with
(a) no exchange, (b) dist.all_reduce (NCCL) of the flat gradient buffer, (c) torch symmetric-memory all-reduce ops.
    python -m torch.distributed.run --nproc-per-node N tools/probe_allreduce.py [steps=50]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from topo4d_b200 import engine, parallel  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    argv = sys.argv
    sys.argv = ["bench.py"]
    a = bench.parse()
    sys.argv = argv
    scene, cams = bench.workload(a)
    views = parallel.shard_views(a.views, rank, world)
    H, W = a.height, a.width
    t = {k: torch.from_numpy(v).to(dev) for k, v in scene.items()}
    cam = torch.tensor(engine.pack_cameras_numpy(cams, (0.0, 0.0, 0.0)), device=dev)[views].contiguous()

    def fwd(cap=None):
        return engine.forward(t["means3D"], t["opacities"], cam, H, W, shs=t["shs"], scales=t["scales"], rotations=t["rotations"],
                              sh_degree=a.sh_degree, check="none" if cap else "sync", cap_instances=cap)
    color, radii, depth, alpha, st = fwd()
    cap = int(st.status().num_instances * 1.1) + 4096
    gimg = (torch.sign(color - 0.5) / (3 * H * W), torch.full_like(depth, 0.1 / (H * W)), torch.full_like(alpha, 0.1 / (H * W)))
    del color, depth, alpha
    n = engine.backward(fwd(cap)[-1], *gimg).flat.numel()
    flat = torch.empty(n, dtype=torch.float32, device=dev)
    modes = {"none": lambda f: None, "nccl": lambda f: dist.all_reduce(f)}
    try:
        import torch.distributed._symmetric_memory as symm
        sflat = symm.empty(n, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(sflat, dist.group.WORLD.group_name)
        modes["symm_two_shot"] = lambda f: torch.ops.symm_mem.two_shot_all_reduce_(sflat, "sum", dist.group.WORLD.group_name)
        modes["symm_one_shot"] = lambda f: torch.ops.symm_mem.one_shot_all_reduce(sflat, "sum", dist.group.WORLD.group_name)
        if getattr(hdl, "multicast_ptr", 0):
            modes["symm_multimem"] = lambda f: torch.ops.symm_mem.multimem_all_reduce_(sflat, "sum", dist.group.WORLD.group_name)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print("symmetric memory unavailable:", repr(e)[:200], flush=True)
        sflat = None
    res = {"world": world, "views_per_rank": len(views), "bytes": n * 4}
    for name, xchg in modes.items():
        buf = sflat if name.startswith("symm") else flat
        try:
            def step():
                *_, s = fwd(cap)
                engine.backward(s, *gimg, flat=buf)
                xchg(buf)
            for _ in range(5):
                step()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[name + "_ms_per_step"] = float(ms.item())
        except Exception as e:  # noqa: BLE001
            res[name + "_error"] = repr(e)[:200]
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
