"""Debug: upstream-style arm vs product path, per 2-D gradient component, one 1080p view of config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity
from tools import upstream_style as US
from topo4d_b200 import engine, synth
opacity = sys.argv[1] if len(sys.argv) > 1 else "topo4d"
dev = torch.device("cuda:0")
scene = synth.head_scene(60000, seed=0, sh_degree=3, opacity=opacity)
cam = synth.ring_cameras(24)[7]
H, W = 1080, 1920
t = {k: torch.tensor(v, device=dev) for k, v in scene.items()}
cam_t = torch.tensor(engine.pack_cameras_numpy([cam], (0, 0, 0)), device=dev)
gen = torch.Generator(device=dev).manual_seed(5)
gC = torch.sign(torch.rand((3, H, W), device=dev, generator=gen) - 0.5) / (3 * H * W)
gD = torch.full((1, H, W), 0.1 / (H * W), device=dev); gA = gD.clone()
color, radii, depth, alpha, view = US.forward(t, cam_t, H, W, 3)
ours = engine.forward(t["means3D"], t["opacities"], cam_t, H, W, shs=t["shs"], scales=t["scales"], rotations=t["rotations"], sh_degree=3)
print("color diff", float((color - ours[0][0]).abs().max()), "pixels>1e-4", int(((color - ours[0][0]).abs().max(0)[0] > 1e-4).sum()))
_, n = engine.flat_layout(60000, 16, True, False)
US.backward(view, gC, gD, gA, torch.empty(n, device=dev))
wv_ptr = view.grad2d_ptr
base = view.ws.data_ptr()
g_us = view.ws[wv_ptr - base: wv_ptr - base + 48 * 60000].view(torch.float32).view(-1, 12).clone().cpu().numpy().astype(np.float64)
engine.backward(ours[4], gC[None], gD[None], gA[None])
g_o = ours[4].view()["grad2d"][0].cpu().numpy().astype(np.float64)
ref = parity.oracle_forward(scene, [cam], H, W, 3, (0, 0, 0))
acc = ref[0]["state"].backward(gC.cpu().numpy(), gD[0].cpu().numpy(), gA[0].cpu().numpy(), return_acc2d=True)["acc2d"]
cols = [0, 1, 2, 3, 4, 5, 8, 9, 10, 6]
names = ["dpix.x", "dpix.y", "dconA", "dconB", "dconC", "dopacity", "dr", "dg", "db", "ddepth"]
for i, (c, nme) in enumerate(zip(cols, names)):
    b = acc[:, i]; s = np.abs(b).max()
    e_us = (np.abs(g_us[:, c] - b) / (np.abs(b) + 1e-3 * s)).max(); e_o = (np.abs(g_o[:, c] - b) / (np.abs(b) + 1e-3 * s)).max()
    print(f"{nme:9s} max|ref| {s:9.3e}  upstream-style vs oracle {e_us:.3e}   ours vs oracle {e_o:.3e}")
# final gradients, 1 view and 3 views accumulated
def rel(a, b):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / (b.abs() + 1e-3 * b.abs().max())).max())
one = torch.empty(n, device=dev)
seg = US.backward(view, gC, gD, gA, one)
gb = engine.backward(ours[4], gC[None], gD[None], gA[None])
print("1 view:", {k: f"{rel(v, getattr(gb, k).reshape(v.shape)):.2e}" for k, v in seg.items()})
cams3 = [synth.ring_cameras(24)[i] for i in (0, 7, 13)]
cam3 = torch.tensor(engine.pack_cameras_numpy(cams3, (0, 0, 0)), device=dev)
tot = torch.zeros(n, device=dev)
for i in range(3):
    c, r, d, a, v = US.forward(t, cam3[i:i + 1].contiguous(), H, W, 3)
    US.backward(v, gC, gD, gA, one)
    tot.add_(one)
o3 = engine.forward(t["means3D"], t["opacities"], cam3, H, W, shs=t["shs"], scales=t["scales"], rotations=t["rotations"], sh_degree=3)
g3 = engine.backward(o3[4], gC[None].expand(3, -1, -1, -1).contiguous(), gD[None].expand(3, -1, -1, -1).contiguous(), gA[None].expand(3, -1, -1, -1).contiguous())
segt = engine.flat_views(tot, 60000, 16, True, False)
print("3 views:", {k: f"{rel(v, getattr(g3, k).reshape(v.shape)):.2e}" for k, v in segt.items()})
