"""Topo4D-shaped optimisation loop on a synthetic multi-view sequence, built only from this repository's pieces
(BASELINE config 5 stand-in: the reference's train.py cannot run here -- pywavefront / trimesh / open3d / skimage /
nvdiffrast and its rasterizer are absent -- so this mirrors its STRUCTURE on synthetic data, function by function):

  initialize_params / initialize_optimizer  (train.py:120-160, 272-297)   parameter dictionary, one Adam group per name
  params2rendervar                          (helpers.py:91-112)           normalize / sigmoid / exp outside the op
  get_loss                                  (train.py:300-327)            Renderer(cam)(**rendervar), camera affine, 0.8 L1 + 0.2 (1-SSIM)
  per-frame loop                            (train.py:640-700)            iterations of loss.backward(); optimizer.step(); pinned rows
  texture bake                              (helpers.py:953-960)          face3d render_colors of a UV mesh -> uint8 texture

View-parallel: with G ranks every step renders G different views (rank r takes view order[step*G + r]), the
gradients are averaged with ONE NCCL all-reduce of a flat buffer, and the replicated FusedAdam steps identically on every
rank; G = 1 is the reference's one-view-per-step loop.

    python tools/train_synthetic.py [--frames 3 --iters 200 --width 512 --height 375 --gaussians 8280] [--json out.json]
    python -m torch.distributed.run --nproc-per-node 8 tools/train_synthetic.py --frames 100 ...
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizer as Renderer  # noqa: E402
from topo4d_b200 import activations, graph, losses, optim, parallel, synth  # noqa: E402
from topo4d_b200.face3d_compat import render as f3d  # noqa: E402


FUSED_ACTIVATIONS = True


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def setup_cameras(n, w, h, dev):
    cams = []
    for c in synth.ring_cameras(n, w=w, h=h, radius=0.6, focal_over_h=1.6):
        w2c = torch.tensor(c.w2c, dtype=torch.float32, device=dev)
        cams.append(Camera(image_height=h, image_width=w, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=torch.zeros(3, device=dev),
                           scale_modifier=1.0, viewmatrix=w2c.unsqueeze(0).transpose(1, 2),
                           projmatrix=torch.tensor(c.projmatrix, device=dev).unsqueeze(0), sh_degree=0,
                           campos=torch.tensor(c.campos, device=dev), prefiltered=False, debug=False))
    return cams


def initialize_params(scene, n_cams, dev):
    t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
    params = {
        "means3D": t["means3D"],
        "rgb_colors": t["colors_precomp"],
        "unnorm_rotations": t["rotations"],
        "logit_opacities": inverse_sigmoid(0.9999 * torch.ones_like(t["opacities"])),       # train.py:142
        "log_scales": torch.log(t["scales"]),
        "cam_m": torch.zeros(n_cams, 3, device=dev),
        "cam_c": torch.zeros(n_cams, 3, device=dev),
    }
    return {k: torch.nn.Parameter(v.float().contiguous()) for k, v in params.items()}


def initialize_optimizer(params, lrs, capturable=False):
    groups = [{"params": [v], "name": k, "lr": lrs[k]} for k, v in params.items()]
    return optim.FusedAdam(groups, lr=0.0, eps=1e-15, capturable=capturable)


def params2rendervar(params):
    if FUSED_ACTIVATIONS:
        return activations.params2rendervar(params)          # same dictionary, one launch each way
    return {"means3D": params["means3D"], "colors_precomp": params["rgb_colors"],
            "rotations": torch.nn.functional.normalize(params["unnorm_rotations"]),
            "opacities": torch.sigmoid(params["logit_opacities"]), "scales": torch.exp(params["log_scales"]),
            "means2D": torch.zeros_like(params["means3D"], requires_grad=True) + 0}


def get_loss(params, cam, cam_id, gt):
    im, radius, _, _ = Renderer(raster_settings=cam)(**params2rendervar(params))
    return losses.image_loss(im, gt, params["cam_m"][cam_id], params["cam_c"][cam_id])


@torch.no_grad()
def render_gt(scene_t, cams, dev):
    rv = {"means3D": scene_t["means3D"], "colors_precomp": scene_t["colors_precomp"],
          "rotations": torch.nn.functional.normalize(scene_t["rotations"]), "opacities": scene_t["opacities"],
          "scales": scene_t["scales"], "means2D": torch.zeros_like(scene_t["means3D"])}
    return [Renderer(raster_settings=c)(**rv)[0].clamp(0, 1) for c in cams]


def deform(means0, t):
    """Low-frequency per-frame deformation of the ground-truth head (a smile-like bulge travelling with t)."""
    m = means0.clone()
    m[:, 0] += 0.004 * torch.sin(40.0 * means0[:, 1] + 0.6 * t)
    m[:, 2] += 0.003 * torch.cos(35.0 * means0[:, 0] - 0.4 * t)
    return m


def psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean().clamp_min(1e-12)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--iters", type=int, default=200, help="optimiser steps per frame")
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--height", type=int, default=375)
    ap.add_argument("--gaussians", type=int, default=8280)
    ap.add_argument("--bake", type=int, default=1024, help="texture size of the per-frame face3d bake (0 = off)")
    ap.add_argument("--graph", action="store_true", help="one CUDA graph per camera for the whole iteration (single rank only)")
    ap.add_argument("--torch-activations", action="store_true", help="helpers.py:91-112 as PyTorch ops instead of the fused kernel")
    ap.add_argument("--json", default="")
    ap.add_argument("--scale-iters", action="store_true", help="G ranks: ceil(iters / G) steps per frame, so the VIEWS SEEN per frame stay "
                    "what the one-view-per-step loop sees (gradients are the mean over the G views of a step)")
    ap.add_argument("--files", default="", help="directory: the ground-truth sequence is written there as JPEG files (rank 0, if absent) in the "
                    "reference's layout <dir>/seq/%%06d/<cam>.jpg and read back through topo4d_b200.frames.FramePrefetcher (GPU decode, "
                    "one frame ahead) instead of being handed over as tensors")
    ap.add_argument("--save-means", default="", help="npy file: final means3D of every frame [frames, N, 3] (rank 0)")
    ap.add_argument("--compare-with", default="", help="npy file written by --save-means of another run (e.g. G = 1): report the distance")
    a = ap.parse_args()
    global FUSED_ACTIVATIONS
    FUSED_ACTIVATIONS = not a.torch_activations
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    gt_scene = synth.head_scene(a.gaussians, seed=0, sh_degree=None, opacity="topo4d")
    gt_t = {k: torch.tensor(v, device=dev) for k, v in gt_scene.items()}
    means0 = gt_t["means3D"].clone()
    cams = setup_cameras(a.views, a.width, a.height, dev)
    # the model starts from the frame-0 geometry with flat grey colours; colours are learnt on frame 0 and then frozen
    # (new_lr["rgb_colors"] = 0.0 after the first timestep, train.py:648), geometry tracks the deformation
    init = dict(gt_scene)
    init["colors_precomp"] = np.full_like(gt_scene["colors_precomp"], 0.5)
    params = initialize_params(init, a.views, dev)
    lrs = {"means3D": 0.0, "rgb_colors": 0.0025 * 4, "unnorm_rotations": 0.001, "logit_opacities": 0.0, "log_scales": 0.001,
           "cam_m": 1e-4, "cam_c": 1e-4}
    if a.graph and world > 1:
        raise SystemExit("--graph is a single-rank option (the view-parallel step contains an NCCL collective)")
    optimizer = initialize_optimizer(params, lrs, capturable=a.graph)
    gt_static = [torch.zeros(3, a.height, a.width, device=dev) for _ in range(a.views)] if a.graph else None
    graphs = {}

    def make_iteration(k):
        def it():
            loss = get_loss(params, cams[k], k, gt_static[k])
            loss.backward()
            optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            return loss
        return it
    # "static" region (back of the head, z > 0.05) is pinned to its initial position after every step (train.py:676)
    static_mask = means0[:, 2] > 0.05
    optimizer.pin(params["means3D"], static_mask, means0)
    uv_v, uv_t, _ = synth.uv_grid_mesh(grid=64, res=max(a.bake, 64), seed=0, extras=False) if a.bake else (None, None, None)

    gen = torch.Generator().manual_seed(0)                   # the same view order on every rank
    iters = -(-a.iters // world) if a.scale_iters else a.iters
    report = {"world": world, "frames": [], "config": vars(a), "steps_per_frame": iters, "views_per_step": world,
              "views_seen_per_frame": iters * world,
              "semantics": "G ranks render G different views per optimiser step (rank r takes order[r] of a fresh permutation), gradients are "
                           "AVERAGED over the G views with one all-reduce, every rank applies the same Adam step; --scale-iters divides the "
                           "steps per frame by G so the views seen per frame match the reference's one-view-per-step loop (train.py:661-673)"}
    prefetch = None
    if a.files:
        from torchvision.io import encode_jpeg
        from topo4d_b200 import frames as fr
        seq_dir = os.path.join(a.files, "seq")
        if rank == 0 and not os.path.exists(os.path.join(seq_dir, "%06d" % a.frames)):
            for t in range(a.frames):                        # write the synthetic sequence once, like a capture session on disk
                gt_t["means3D"] = deform(means0, float(t))
                d = os.path.join(seq_dir, "%06d" % (t + 1))
                os.makedirs(d, exist_ok=True)
                for k, im in enumerate(render_gt(gt_t, cams, dev)):
                    open(os.path.join(d, "cam%02d.jpg" % k), "wb").write(
                        encode_jpeg((im * 255).round().to(torch.uint8).cpu(), quality=95).numpy().tobytes())
        if world > 1:
            dist.barrier()
        prefetch = fr.FramePrefetcher(lambda t: fr.list_frame_files(a.files, "seq", t + 1), [0] * a.views, dev, num_frames=a.frames)
    saved = []
    t_start = time.perf_counter()
    steps_total = 0
    for t in range(a.frames):
        gt_t["means3D"] = deform(means0, float(t))
        gts = prefetch.get(t) if prefetch else render_gt(gt_t, cams, dev)
        if a.graph:
            for k in range(a.views):
                gt_static[k].copy_(gts[k])                   # graphs read static buffers: new frame, same storage
        if t == 1:
            for g in optimizer.param_groups:                 # update_optimizer (helpers.py:801-804)
                g["lr"] = {"rgb_colors": 0.0, "means3D": 0.00016}.get(g["name"], g["lr"])
            optimizer.sync_hyperparams()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        first = last = None
        for i in range(iters):
            order = torch.randperm(a.views, generator=gen).tolist()
            cam_id = order[rank % a.views] if world > 1 else order[0]
            if a.graph:
                if cam_id not in graphs:                     # first visit: 2 eager steps, then the capture (not executed)
                    graphs[cam_id] = graph.capture(make_iteration(cam_id), warmup=2, capacity_headroom=3.0)
                    steps_total += 1
                    loss = graphs[cam_id].outputs
                else:
                    loss = graphs[cam_id].replay()
                if i == 0:
                    first = loss.detach().clone()
                last = loss
                steps_total += 1
                continue
            loss = get_loss(params, cams[cam_id], cam_id, gts[cam_id])
            loss.backward()
            parallel.allreduce_param_grads_(params, average=True)    # ONE all-reduce of a flat buffer per step (no-op at 1 rank)
            optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            if i == 0:
                first = loss.detach()
            last = loss.detach()
            steps_total += 1
        e1.record()
        torch.cuda.synchronize()
        for g in graphs.values():
            g.check()                                        # did any replay outgrow its captured workspace?
        with torch.no_grad():
            ims = [Renderer(raster_settings=cams[k])(**params2rendervar(params))[0] for k in (0, a.views // 2)]
            p = float(np.mean([psnr(torch.exp(params["cam_m"][k])[:, None, None] * im + params["cam_c"][k][:, None, None], gts[k])
                               for im, k in zip(ims, (0, a.views // 2))]))
            geo = float((params["means3D"][~static_mask] - gt_t["means3D"][~static_mask]).norm(dim=1).mean())
            pinned = float((params["means3D"][static_mask] - means0[static_mask]).abs().max())
        frame = {"frame": t, "loss_first": float(first), "loss_last": float(last), "psnr_db": p, "mean_vertex_err_m": geo,
                 "pinned_rows_max_dev": pinned, "ms_per_step": e0.elapsed_time(e1) / iters}
        if a.save_means or a.compare_with:
            saved.append(params["means3D"].detach().cpu().numpy().copy())
        if a.bake and rank == 0:                             # texture bake of the learnt colours (helpers.py:953-960)
            tb = time.perf_counter()
            col = params["rgb_colors"].detach().clamp(0, 1).cpu().numpy().astype(np.float64)
            vc = col[np.arange(uv_v.shape[0]) % col.shape[0]]
            tex = f3d.render_colors_u8(uv_v, uv_t, vc, a.bake, a.bake)
            frame["bake_ms"] = (time.perf_counter() - tb) * 1e3
            frame["bake_mean_u8"] = float(tex.mean())
        report["frames"].append(frame)
        if rank == 0:
            print(json.dumps(frame), flush=True)
    report["steps_per_s"] = steps_total / (time.perf_counter() - t_start)
    report["views_per_s"] = report["steps_per_s"] * world
    report["seconds_total"] = time.perf_counter() - t_start
    report["final_mean_vertex_err_m"] = report["frames"][-1]["mean_vertex_err_m"]
    report["mean_over_frames_vertex_err_m"] = float(np.mean([f["mean_vertex_err_m"] for f in report["frames"][1:] or report["frames"]]))
    if rank == 0 and a.save_means:
        np.save(a.save_means, np.stack(saved))
    if rank == 0 and a.compare_with and os.path.exists(a.compare_with):
        other = np.load(a.compare_with)
        n = min(len(other), len(saved))
        dist_m = np.linalg.norm(np.stack(saved)[:n] - other[:n], axis=-1)                # [frames, N]
        report["vs_other_run"] = {"file": a.compare_with, "frames_compared": int(n), "mean_vertex_distance_m": float(dist_m.mean()),
                                  "max_vertex_distance_m": float(dist_m.max()), "final_frame_mean_m": float(dist_m[-1].mean()),
                                  "deformation_amplitude_m": 0.004,
                                  "note": "distance between the meshes of this run and of the other run, per frame and vertex"}
    if rank == 0:
        print(json.dumps({k: v for k, v in report.items() if k != "frames"}), flush=True)
        if a.json:
            json.dump(report, open(a.json, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
