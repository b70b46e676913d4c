"""CPU, world_size 2, gloo: the view-parallel host logic (sharding + one all-reduce of the flat gradient
buffer) reproduces the single-process sum over all views.  Per-view gradients come from the oracle here
(no GPU in this container); on the GPU box the same helpers run over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from topo4d_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_grads(scene, cam, H, W, seed):
    from oracle import gs_oracle
    rng = np.random.default_rng(seed)
    _, _, _, _, st = gs_oracle.forward(scene["means3D"], scene["opacities"], colors_precomp=scene["colors_precomp"],
                                       scales=scene["scales"], rotations=scene["rotations"], image_height=H, image_width=W,
                                       tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.zeros(3, np.float32),
                                       viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos)
    g = st.backward(rng.normal(size=(3, H, W)).astype(np.float32))
    return np.concatenate([g[k].reshape(-1) for k in ("means3D", "means2D", "colors_precomp", "opacities", "scales", "rotations")])


def _worker(rank, world, port, n_views, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = synth.random_scene(300, seed=0)
    cams = synth.ring_cameras(n_views, w=64, h=48, radius=4.0, focal_over_h=1.2)
    mine = parallel.shard_views(n_views, rank, world)
    flat = torch.zeros(300 * 17, dtype=torch.float32)
    for v in mine:
        flat += torch.from_numpy(_view_grads(scene, cams[v], 48, 64, seed=v))
    parallel.allreduce_flat_(flat)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), flat.numpy())
    dist.destroy_process_group()


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in parallel.shard_views(24, r, world))
        assert seen == list(range(24))
        assert {len(parallel.shard_views(24, r, world)) for r in range(world)} == {24 // world}
    assert parallel.shard_views(5, 1, 2) == [1, 3]


def test_allreduce_is_identity_without_process_group():
    x = torch.arange(8, dtype=torch.float32)
    assert parallel.allreduce_flat_(x.clone()).equal(x)


def test_two_rank_view_parallel_matches_single_process(tmp_path):
    n_views, world = 4, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_views, str(tmp_path)), nprocs=world, join=True)
    scene = synth.random_scene(300, seed=0)
    cams = synth.ring_cameras(n_views, w=64, h=48, radius=4.0, focal_over_h=1.2)
    ref = sum(_view_grads(scene, cams[v], 48, 64, seed=v).astype(np.float64) for v in range(n_views))
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)                      # every rank ends with the same reduced buffer
    assert np.abs(ref).max() > 0
    np.testing.assert_allclose(r0, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    params = {"means3D": torch.nn.Parameter(torch.zeros(50, 3)), "unnorm_rotations": torch.nn.Parameter(torch.zeros(50, 4)),
              "frozen": torch.nn.Parameter(torch.zeros(7)), "cam_m": torch.nn.Parameter(torch.zeros(24, 3))}
    for k, p in params.items():
        if k != "frozen":                              # no gradient on any rank: skipped, buffers still line up
            p.grad = torch.randn(p.shape, generator=g)
    parallel.allreduce_param_grads_(params, average=True)
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), **{k: p.grad.numpy() for k, p in params.items() if p.grad is not None})
    dist.destroy_process_group()


def test_two_rank_parameter_gradient_exchange(tmp_path):
    """The training loop's exchange (tools/train_synthetic.py): every rank ends with the MEAN of the ranks' gradients,
    parameter by parameter, from one flat all-reduce; without a process group it is the identity."""
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "g0.npz"), np.load(tmp_path / "g1.npz")
    assert sorted(a.files) == ["cam_m", "means3D", "unnorm_rotations"]
    for k in a.files:
        np.testing.assert_array_equal(a[k], b[k])          # every rank ends with the same gradients
    # regenerate what each rank drew (same generator protocol as the worker) and compare with the mean
    draws = []
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        draws.append({k: torch.randn(s, generator=g).numpy() for k, s in (("means3D", (50, 3)), ("unnorm_rotations", (50, 4)), ("cam_m", (24, 3)))})
    for k in a.files:
        np.testing.assert_allclose(a[k], 0.5 * (draws[0][k] + draws[1][k]), rtol=1e-6, atol=1e-7)
    p = {"x": torch.nn.Parameter(torch.zeros(3))}
    p["x"].grad = torch.ones(3)
    assert parallel.allreduce_param_grads_(p)["x"].grad.equal(torch.ones(3))
