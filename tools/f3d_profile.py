"""Tiny driver for ncu: warm-up + one 8192^2 bake (device tensors only): the fused bake path and the in-place API path.
    ncu --set full --clock-control none -k regex:f3d -c 12 -o gpurun_out/f3d python tools/f3d_profile.py [grid]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topo4d_b200 import synth
from topo4d_b200.face3d_compat import mesh_core_cython as mcc
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 245
res = 8192
v, t, c = synth.uv_grid_mesh(grid=grid, res=res, seed=0)
dev = torch.device("cuda:0")
d_v = torch.tensor(v, dtype=torch.float32, device=dev); d_t = torch.tensor(t, dtype=torch.int32, device=dev); d_c = torch.tensor(c, dtype=torch.float32, device=dev)
for _ in range(2):
    out = mcc.bake_colors_device(d_v, d_t, d_c, res, res, 3)
torch.cuda.synchronize()
