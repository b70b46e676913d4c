"""Tiny driver for ncu: one warm-up + one timed 8192^2 bake (device tensors only)."""
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
ws = None
for _ in range(2):
    img = torch.zeros((res, res, 3), device=dev); dep = torch.full((res, res), -999999.0, device=dev)
    ws = mcc.render_colors_device(img, d_v, d_t, d_c, dep, res, res, 3, ws)
torch.cuda.synchronize()
